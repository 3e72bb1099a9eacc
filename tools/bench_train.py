#!/usr/bin/env python
"""BASELINE.json configs[3]: decoder training step, bf16, synthetic 16x64x64 videos, data-parallel with an NCCL gradient
all-reduce at N GPUs.

    python tools/bench_train.py [--batch 8] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

One step = VideoDecoderTrainer.forward (q_sample, Unet3D forward, loss, backward, gradient all-reduce) + update
(clip_grad_norm_ 0.5, AdamW).  Timed with CUDA events after warm-up, barrier + synchronize on both sides, max over ranks;
prints ONE JSON line with samples/s (all ranks), the per-phase breakdown of the last step and the all-reduce exposed time.
Algorithmic work: 3 x 291.8 = 875 GFLOP per sample."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNET_KW = dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True)
FRAMES, SIZE = 16, 64
GFLOP_PER_SAMPLE = 3 * 291.8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--overlap", action="store_true", help="all-reduce in buckets launched under the backward instead of once after it")
    args = ap.parse_args()
    import torch.distributed as dist
    from dalle2_video_b200 import Unet3D, VideoDecoder, VideoDecoderTrainer
    from dalle2_video_b200 import _lib as L
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    unet = Unet3D(**UNET_KW)
    torch.nn.init.normal_(unet.to_out.weight, std=0.02)
    dec = VideoDecoder(unet, frame_size=args.size, frame_numbers=(args.frames,), timesteps=1000, learned_variance=False).to(dev)
    trainer = VideoDecoderTrainer(dec, lr=3e-4, wd=1e-2, max_grad_norm=0.5, overlap_allreduce=args.overlap)
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    video = torch.rand(B, 3, args.frames, args.size, args.size, device=dev, generator=g)
    emb = torch.randn(B, 512, device=dev, generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        losses.append(trainer(video, video_embed=emb))
        trainer.update()
    barrier()
    c0 = L.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses.append(trainer(video, video_embed=emb))
        trainer.update()
    e1.record()
    barrier()
    elapsed = e0.elapsed_time(e1) * 1e-3
    launches = L.launch_count - c0

    # phase breakdown of one more step (device time per phase, eager)
    plan = dec.unets[0].train_plan(B, args.frames, args.size, args.size, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    torch.cuda.synchronize()
    trainer._overlap_live = False   # this extra step reduces in one piece after the backward so the phases separate
    ev[0].record()
    loss = dec(video, video_embed=emb)
    ev[1].record()
    loss.backward()
    ev[2].record()
    if world > 1:
        dist.all_reduce(trainer.flat_g)
        trainer.flat_g.div_(world)
    ev[3].record()
    trainer.update()
    ev[4].record()
    torch.cuda.synchronize()
    phases = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(("forward_ms", "backward_ms", "allreduce_ms", "update_ms"))}
    if world > 1:
        t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t[0])
    if rank == 0:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = json.load(open(pk))["bf16_tflops_sustained"] if os.path.isfile(pk) else 1400.0
        value = world * B * args.steps / elapsed
        print(json.dumps({
            "metric": "decoder_train_samples_per_sec_16x64x64", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "config4: decoder training step (p_losses + backward + clip + AdamW), Unet3D dim64 "
                                   f"mults(1,2,4,8), 3x{args.frames}x{args.size}x{args.size}", "batch_per_gpu": B,
                       "parallelism": f"data-parallel x{world}, NCCL all-reduce of the flat fp32 gradient once per step, "
                                      + ("in buckets launched under the backward" if trainer._buckets is not None
                                         else "one call after the backward")},
            "allreduce": None if trainer._buckets is None else {
                "backward_segments": len(plan.segments),
                "segment_launches": [hi - lo for lo, hi, _ in plan.segments],
                "buckets_mb_by_ready_segment": [[round((hi - lo) * 4 / 2 ** 20, 2), seg] for lo, hi, seg in trainer._buckets.buckets]},
            "model_tflops": value * GFLOP_PER_SAMPLE / 1e3 * (args.frames * args.size ** 2) / (FRAMES * SIZE ** 2),
            "frac_of_sustained_peak_per_gpu": value * GFLOP_PER_SAMPLE / 1e3 * (args.frames * args.size ** 2) / (FRAMES * SIZE ** 2) / world / peak,
            "phases_last_step": phases, "grad_bytes": int(trainer.flat_g.numel() * 4),
            "gpu_launches": int(launches), "launches_per_step": launches // args.steps,
            "fwd_launches": len(plan.fwd), "bwd_launches": len(plan.bwd),
            "loss_first": losses[0], "loss_last": losses[-1],
            "mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
