#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--dtype bf16|fp32] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (configs[1], "config2"): Unet3D(dim=64, dim_mults=(1,2,4,8), video_embed_dim=512,
cond_on_video_embeds=True) denoising 3x16x64x64 videos inside the 1000-step cosine DDPM loop, bf16 compute,
fp32 sampler state, random-init weights (to_out re-randomised: it is zero-initialised in the reference),
synthetic N(0,1) video embeddings.  One "step" = one p_sample step (Unet3D forward + fused posterior
update, incl. the per-step noise draw) over the rank's batch of B videos; ranks hold disjoint batches (weak
scaling, no data-path collective).  value = total sample-steps/s over all ranks; the JSON also carries
frames/s (16 frames per sample-step) and videos/s at 1000 steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_KW = dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True)
FRAMES, SIZE, TIMESTEPS = 16, 64, 1000
GFLOP_PER_SAMPLE_STEP = 291.8  # reference formulation, FlopCounterMode on the reference module
METRIC, UNIT = "unet3d_denoise_sample_steps_per_sec_16x64x64", "sample-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [s.strip() for s in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def build_model(device, dtype, seed=0):
    import torch
    from dalle2_video_b200 import Unet3D, VideoDecoder
    torch.manual_seed(seed)
    unet = Unet3D(**UNET_KW)
    torch.nn.init.normal_(unet.to_out.weight, std=0.02)  # zero-initialised in the reference
    dec = VideoDecoder(unet, frame_size=SIZE, frame_numbers=(FRAMES,), timesteps=TIMESTEPS, learned_variance=False)
    dec = dec.to(device)
    dec.unets[0].compute_dtype = dtype
    return dec


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step_time(n_steps, warmup, threads=None):
    """The reference's algorithm for this path on the host cores: the oracle restatement (oracle/unet3d.py; the
    reference itself cannot travel to the GPU box) — one denoise step = Unet3D forward + posterior update, B=1."""
    import torch
    from oracle.primitives import NoiseScheduler
    from oracle.unet3d import UnetConfig, ddpm_posterior_step, unet3d_forward
    from oracle.weights import synth_state_dict
    from dalle2_video_b200 import Unet3D
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    shapes = {k: list(v.shape) for k, v in Unet3D(**UNET_KW).state_dict().items()}
    sd = synth_state_dict(shapes, seed=0)
    cfg = UnetConfig(**UNET_KW)
    sched = NoiseScheduler(beta_schedule="cosine", timesteps=TIMESTEPS, loss_type="l2")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, FRAMES, SIZE, SIZE, generator=g)
    emb = torch.randn(1, 512, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + n_steps):
            t = TIMESTEPS - 1 - i
            t0 = time.perf_counter()
            eps = unet3d_forward(sd, cfg, x, torch.tensor([t]), video_embed=emb)
            x, _ = ddpm_posterior_step(sched, x, t, eps, torch.randn(x.shape, generator=g))
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, threads = cpu_reference_step_time(args.steps, args.warmup)
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: Unet3D dim64 mults(1,2,4,8) 3x16x64x64, 1000-step cosine DDPM, one p_sample step",
                       "batch_per_step": 1},
            "frames_per_sec": val * FRAMES,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} denoise steps at batch 1 after {args.warmup} warm-up, oracle restatement "
                                       "of the reference Unet3D + DDPM update, fp32, torch CPU"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def conv_flops(p):
    taps = 4 if p.mode == 1 else p.kh * p.kw
    ho, wo = (p.H // 2, p.W // 2) if p.mode == 1 else (p.H, p.W)
    return 2.0 * p.n_img * ho * wo * p.Cout * taps * (p.C0 + p.C1)


def profile_conv_family(eng, iters=3):
    """CUDA-event time of every dv_conv_igemm launch of one step (eager, on the launch stream) and the
    algorithmic FLOPs they perform; returns (flops_per_step, seconds_per_step, launches, tc_launches)."""
    import torch
    from dalle2_video_b200 import _lib as L
    stream = torch.cuda.current_stream().cuda_stream
    convs = [(i, p) for i, (n, p) in enumerate(eng.ops) if n == "dv_conv_igemm"]
    flops = sum(conv_flops(p) for _, p in convs)
    tc = sum(1 for _, p in convs if L.lib().dv_conv_uses_tcgen05(p))
    total = 0.0
    # the single dominant launch shape: the 3x3 64->64 convolution at full resolution (14 of the 88 launches, 35 % of the
    # family's time) — its per-launch duration is what roofline.traffic (one ncu --set full capture) is quoted against
    def is_dom(p):
        return p.kh == 3 and p.H == SIZE and p.C0 == 64 and p.C1 == 0 and p.Cout == 64 and p.mode == 0
    dom_n = sum(1 for _, p in convs if is_dom(p))
    dom_flops = conv_flops(next(p for _, p in convs if is_dom(p))) if dom_n else 0.0
    dom_sec = 0.0
    for it in range(iters + 1):
        evs = []
        eng.stats.zero_()
        for name, p in eng.ops:
            if name == "dv_conv_igemm":
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                L.call(name, p, stream)
                b.record()
                evs.append((a, b, is_dom(p)))
            else:
                L.call(name, p, stream)
        torch.cuda.synchronize()
        if it > 0:
            total += sum(a.elapsed_time(b) for a, b, _ in evs) * 1e-3
            dom_sec += sum(a.elapsed_time(b) for a, b, d in evs if d) * 1e-3
    return flops, total / iters, len(convs), tc, dom_sec / iters / max(dom_n, 1), dom_flops, dom_n


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from dalle2_video_b200 import _lib as L
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B = args.batch
    dec = build_model(dev, dtype)
    unet, sched = dec.unets[0], dec.noise_schedulers[0]
    shape = (B, 3, FRAMES, SIZE, SIZE)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    emb = torch.randn(B, 512, device=dev, generator=gen)

    # ---- device-resident loop: the sampler's own step (graph replay), state and conditioning already in HBM
    eng = unet.engine(B, FRAMES, SIZE, SIZE, dtype, dev, 0)
    keep = torch.ones(B, dtype=torch.uint8, device=dev)
    eng.set_conditioning(emb, None, keep, keep)
    eng.set_time_source(True)
    x = eng.x_in
    x.copy_(torch.randn(shape, device=dev, generator=gen))
    noise = torch.empty(shape, device=dev)
    coef = sched.coef_table().to(dev)
    params = L.DdpmStepParams(x_t=x.data_ptr(), eps=eng.eps.data_ptr(), eps_null=None, noise=noise.data_ptr(),
                              out=x.data_ptr(), unnormalize_out=None, coef=coef.data_ptr(), t_dev=eng.t_dev.data_ptr(),
                              t_host=0, t_dev_decrement=1, num_timesteps=TIMESTEPS, n=x.numel(), cond_scale=1.0,
                              clip_denoised=1, eps_dtype=L.DV_F32)

    def step_body():
        eng.run()
        L.call("dv_ddpm_step", params, torch.cuda.current_stream().cuda_stream)

    eng.t_dev.fill_(TIMESTEPS - 1)
    c0 = L.launch_count
    step_body()                       # eager warm-up (module load, function attributes)
    launches_per_step = L.launch_count - c0 + 1  # + the statistics memset
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            step_body()
    torch.cuda.current_stream().wait_stream(side)

    def one_step():
        noise.normal_(generator=gen)  # the per-step randn_like of p_sample (:1657)
        graph.replay()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.t_dev.fill_(TIMESTEPS - 1)
    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    sampler.stop_flag.set()
    elapsed = e0.elapsed_time(e1) * 1e-3
    assert bool(torch.isfinite(x).all()), "non-finite sampler state"

    # ---- end to end through the public API with HOST buffers: VideoDecoder.p_sample per step
    hx = torch.randn(shape).pin_memory()
    hemb = emb.cpu().pin_memory()
    ht = torch.full((B,), TIMESTEPS - 1, dtype=torch.long).pin_memory()
    hout = torch.empty(shape).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(i):
        xd = hx.to(dev, non_blocking=True)
        ed = hemb.to(dev, non_blocking=True)
        td = (ht - i).to(dev, non_blocking=True)
        out, _ = dec.p_sample(unet, xd, td, ed, sched)
        hout.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_elapsed = time.perf_counter() - t0

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([elapsed, e2e_elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, e2e_elapsed = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        flops, conv_sec, n_conv, n_tc, dom_sec, dom_flops, dom_n = profile_conv_family(eng)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get(f"conv_halo_3x3_64x64_C64_C64_B{B}")
        achieved = flops / conv_sec / 1e12
        value = world * B * args.steps / elapsed
        e2e_val = world * B * e2e_steps / e2e_elapsed
        ws_gb = eng.pool.total / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sec, threads = cpu_reference_step_time(3, 1)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "3 denoise steps at batch 1 after 1 warm-up (oracle restatement of the reference Unet3D + "
                             "DDPM update, fp32, torch CPU, all host threads)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "config2: Unet3D dim64 mults(1,2,4,8) cond_on_video_embeds, 3x16x64x64, 1000-step cosine "
                                   "DDPM, one p_sample step (forward + fused posterior update + noise draw)",
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"batch-sharded x{world}, no collective",
                       "l2": f"activation working set {ws_gb:.1f} GB per step >> 126 MB L2 (no flush needed)",
                       "cuda_graph": True},
            "frames_per_sec": value * FRAMES, "videos_per_sec_at_1000_steps": value / TIMESTEPS,
            "model_tflops": value * GFLOP_PER_SAMPLE_STEP / 1e3,
            "model_frac_of_sustained_peak_per_gpu": value * GFLOP_PER_SAMPLE_STEP / 1e3 / world / pk["bf16_sustained"],
            "e2e": {"value": e2e_val, "unit": UNIT,
                    "h2d_bytes_per_step": int(hx.numel() * 4 + hemb.numel() * 4 + ht.numel() * 8),
                    "d2h_bytes_per_step": int(hout.numel() * 4), "steps": e2e_steps,
                    "api": "VideoDecoder.p_sample(unet, x, t, video_embed, scheduler) with pinned host x/t/embed in, x_(t-1) out"},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "dv_conv_igemm (all launches of one step)", "achieved": achieved,
                         "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                         "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside the step)",
                         "algorithmic_gflop_per_step": flops / 1e9, "conv_ms_per_step": conv_sec * 1e3,
                         "launches": n_conv, "tcgen05_launches": n_tc,
                         "dominant_launch": {"kernel": "conv_halo_kernel 3x3 64->64 @64x64 (dv_conv_igemm)", "launches_per_step": dom_n,
                                             "algorithmic_gflop": dom_flops / 1e9, "avg_ms": dom_sec * 1e3,
                                             "achieved": dom_flops / dom_sec / 1e12 if dom_sec else None,
                                             "frac": dom_flops / dom_sec / 1e12 / pk["bf16_sustained"] if dom_sec else None,
                                             "algorithmic_bytes": 2.0 * 2 * B * FRAMES * SIZE * SIZE * 64},
                         "traffic": traffic,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant launch shape, "
                                         "one ncu --set full capture (profiles/ncu_traffic.json names the summaries); null for batches "
                                         "that were not captured"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="videos per GPU (SURVEY.md §8d sweeps 1..32; measured 32 -> 2095, 48 -> 2146, 64 -> 2160 sample-steps/s: the 8x8 / 16x16 layers keep filling)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
