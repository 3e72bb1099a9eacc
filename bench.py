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
def _synth_weights():
    from oracle.weights import synth_state_dict
    from dalle2_video_b200 import Unet3D
    shapes = {k: list(v.shape) for k, v in Unet3D(**UNET_KW).state_dict().items()}
    return synth_state_dict(shapes, seed=0)


def cpu_reference_step_time(n_steps, warmup, threads=None):
    """The reference's CPU implementation of the path on the host cores, one denoise step at B=1 in fp32 with all host
    threads: `VideoDecoder.p_sample` of the UNMODIFIED reference package when it is present (baseline/_ref, installed
    from /root/reference — its Unet3D / VideoDecoder code over oracle/primitives.py's restatement of the un-installable
    dalle2-pytorch symbols; kind "reference"), else the oracle restatement of the same step (oracle/unet3d.py; kind
    "port").  Returns (seconds per step, threads, kind, description)."""
    import torch
    from oracle import reference
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = _synth_weights()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, FRAMES, SIZE, SIZE, generator=g)
    emb = torch.randn(1, 512, generator=g)
    times = []
    if reference.available():
        ref = reference.load()
        dec = ref.VideoDecoder(ref.Unet3D(**UNET_KW), frame_size=SIZE, frame_numbers=(FRAMES,), timesteps=TIMESTEPS,
                               learned_variance=False)
        dec.unets[0].load_state_dict(sd, strict=True)   # after construction: VideoDecoder re-instantiates its unets (F4)
        dec.eval()
        unet, sched = dec.unets[0], dec.noise_schedulers[0]
        torch.manual_seed(0)
        with torch.no_grad():
            for i in range(warmup + n_steps):
                t = torch.full((1,), TIMESTEPS - 1 - i, dtype=torch.long)
                t0 = time.perf_counter()
                x, _ = dec.p_sample(unet, x, t, video_embed=emb, noise_scheduler=sched)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
        kind = "reference"
        what = (f"the reference's own VideoDecoder.p_sample (unmodified package from {reference.REFERENCE_ROOT}, "
                "dalle2-pytorch symbols restated in oracle/primitives.py)")
    else:
        from oracle.primitives import NoiseScheduler
        from oracle.unet3d import UnetConfig, ddpm_posterior_step, unet3d_forward
        cfg = UnetConfig(**UNET_KW)
        sched = NoiseScheduler(beta_schedule="cosine", timesteps=TIMESTEPS, loss_type="l2")
        with torch.no_grad():
            for i in range(warmup + n_steps):
                t = TIMESTEPS - 1 - i
                t0 = time.perf_counter()
                eps = unet3d_forward(sd, cfg, x, torch.tensor([t]), video_embed=emb)
                x, _ = ddpm_posterior_step(sched, x, t, eps, torch.randn(x.shape, generator=g))
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
        kind, what = "port", "oracle restatement of the reference Unet3D + DDPM update (oracle/unet3d.py)"
    return sum(times) / len(times), threads, kind, what


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, threads, kind, what = cpu_reference_step_time(args.steps, args.warmup)
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: Unet3D dim64 mults(1,2,4,8) 3x16x64x64, 1000-step cosine DDPM, one p_sample step",
                       "batch_per_step": 1},
            "frames_per_sec": val * FRAMES,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{args.steps} denoise steps at batch 1 after {args.warmup} warm-up: {what}, fp32, "
                                       "torch CPU, all host threads"},
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
    allc = [(i, p) for i, (n, p) in enumerate(eng.ops) if n == "dv_conv_igemm"]
    # The 1x1 res_conv launches whose epilogue also applies block2's GroupNorm + SiLU to the residual operand (add_*) move
    # four tensor-sized streams for K = 128..384 MACs per output: they are HBM-bound by construction and replace a
    # dv_gn_apply pass, so they are timed and reported against the HBM roofline (roofline.hbm_fused_residual), not inside
    # the tensor-bound family
    fused = [(i, p) for i, p in allc if p.add_stats]
    convs = [(i, p) for i, p in allc if not p.add_stats]
    esz = 2
    fused_bytes = sum(esz * p.n_img * p.H * p.W * (p.C0 + p.C1 + 2 * p.Cout) for _, p in fused)
    fused_ids = {i for i, _ in fused}
    fused_sec = 0.0
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
        fevs = []
        for idx, (name, p) in enumerate(eng.ops):
            if name == "dv_conv_igemm":
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                L.call(name, p, stream)
                b.record()
                if idx in fused_ids:
                    fevs.append((a, b))
                else:
                    evs.append((a, b, is_dom(p)))
            else:
                L.call(name, p, stream)
        torch.cuda.synchronize()
        if it > 0:
            total += sum(a.elapsed_time(b) for a, b, _ in evs) * 1e-3
            dom_sec += sum(a.elapsed_time(b) for a, b, d in evs if d) * 1e-3
            fused_sec += sum(a.elapsed_time(b) for a, b in fevs) * 1e-3
    return (flops, total / iters, len(convs), tc, dom_sec / iters / max(dom_n, 1), dom_flops, dom_n,
            (len(fused), fused_sec / iters, fused_bytes))


def gpu_eager_baseline(dev, B, steps=3):
    """"the real kernel to beat" — the reference formulation as plain torch eager ops on the SAME
    GPU (the oracle restatement moved to cuda: every conv / GroupNorm / attention is a cuDNN / cuBLAS / ATen kernel),
    bf16 autocast, the bench's batch.  A reported baseline leg like cpu_baseline: measurement only, after the timed
    regions of the product path, never part of them."""
    import torch
    from oracle.primitives import NoiseScheduler
    from oracle.unet3d import UnetConfig, ddpm_posterior_step, unet3d_forward
    sd = {k: v.to(dev) for k, v in _synth_weights().items()}
    cfg = UnetConfig(**UNET_KW)
    sched = NoiseScheduler(beta_schedule="cosine", timesteps=TIMESTEPS, loss_type="l2").to(dev)
    while B >= 1:
        try:
            x = torch.randn(B, 3, FRAMES, SIZE, SIZE, device=dev)
            emb = torch.randn(B, 512, device=dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                for i in range(1 + steps):
                    if i == 1:
                        torch.cuda.synchronize()
                        ev[0].record()
                    t = TIMESTEPS - 1 - i
                    eps = unet3d_forward(sd, cfg, x, torch.full((B,), t, device=dev), video_embed=emb)
                    x, _ = ddpm_posterior_step(sched, x, t, eps.float(), torch.randn_like(x))
                ev[1].record()
            torch.cuda.synchronize()
            sec = ev[0].elapsed_time(ev[1]) * 1e-3 / steps
            return {"value": B / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "batch": B, "steps": steps,
                    "dtype": "bf16 autocast", "kind": "torch eager (cuDNN / cuBLAS / ATen) on the same GPU, oracle "
                    "restatement of the reference Unet3D + DDPM update"}
        except torch.OutOfMemoryError:
            torch.cuda.empty_cache()
            B //= 2
    return None


def train_block(dev, world, rank, B, steps, warmup):
    """BASELINE.json configs[3] / SURVEY.md row a16 + section 8e: the decoder training step (q_sample, Unet3D forward,
    loss, native backward, NCCL all-reduce of the flat fp32 gradient — the ONE collective of the path — clip 0.5 + AdamW),
    bf16 activations, synthetic 16x64x64 videos, B per GPU, data-parallel over the ranks.  Timed like the headline
    (events, barrier + synchronize on both sides, max over ranks).  The exposed all-reduce time is the difference to the
    same steps with the collective skipped."""
    import torch
    import torch.distributed as dist
    from dalle2_video_b200 import Unet3D, VideoDecoder, VideoDecoderTrainer
    from dalle2_video_b200 import _lib as L
    from dalle2_video_b200 import train as T
    torch.manual_seed(1 + rank)                       # replicas are synchronised by the trainer, not by the seed
    unet = Unet3D(**UNET_KW)
    torch.nn.init.normal_(unet.to_out.weight, std=0.02)
    dec = VideoDecoder(unet, frame_size=SIZE, frame_numbers=(FRAMES,), timesteps=TIMESTEPS, learned_variance=False).to(dev)
    trainer = VideoDecoderTrainer(dec, lr=3e-4, wd=1e-2, max_grad_norm=0.5)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    video = torch.rand(B, 3, FRAMES, SIZE, SIZE, device=dev, generator=g)
    emb = torch.randn(B, 512, device=dev, generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n):
        barrier()
        c0 = L.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            losses.append(trainer(video, video_embed=emb))
            trainer.update()
        e1.record()
        barrier()
        sec = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t[0])
        return sec, L.launch_count - c0

    losses = []
    for _ in range(max(warmup, 3)):                   # eager, graph capture, first replay
        losses.append(trainer(video, video_embed=emb))
        trainer.update()
    sec, launches = timed(steps)
    exposed = None
    if world > 1:
        real = T.allreduce_mean_
        T.allreduce_mean_ = lambda flat: flat         # same steps without the collective
        try:
            sec_local, _ = timed(steps)
        finally:
            T.allreduce_mean_ = real
        trainer.sync_replicas()
        exposed = (sec - sec_local) / steps * 1e3
    value = world * B * steps / sec
    pk = peaks()
    out = {"metric": "decoder_train_samples_per_sec_16x64x64", "value": value, "unit": "samples/s",
           "ms_per_step": sec / steps * 1e3, "steps": steps, "batch_per_gpu": B, "n_gpus": world, "dtype": "bf16",
           "scaling": "weak", "higher_is_better": True,
           "collective": None if world == 1 else "NCCL all-reduce (sum, then divide by the world size) of the flat fp32 "
                                                 "gradient buffer, one call per step after the backward",
           "allreduce_exposed_ms": exposed, "grad_bytes": int(trainer.flat_g.numel() * 4),
           "grad_bytes_if_bf16": int(trainer.flat_g.numel() * 2),
           "model_tflops": value * 3 * GFLOP_PER_SAMPLE_STEP / 1e3,
           "frac_of_sustained_peak_per_gpu": value * 3 * GFLOP_PER_SAMPLE_STEP / 1e3 / world / pk["bf16_sustained"],
           "gpu_launches": int(launches), "loss_first": losses[0], "loss_last": losses[-1],
           "mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
           "workload": "config4: VideoDecoderTrainer.forward (p_losses + backward + gradient all-reduce) + update (clip "
                       "0.5, AdamW), Unet3D dim64 mults(1,2,4,8), 3x16x64x64, 875 GFLOP per sample (3 x 291.8)"}
    del trainer, dec, unet
    torch.cuda.empty_cache()
    return out


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
    # 40 steps (1 s): the wall-clock region contains the pipeline's fill (first H2D) and drain (last D2H), 2 x ~2 ms, and
    # any host-side hiccup; over 10 steps one hiccup moved the figure by 10 %
    e2e_steps = max(3, min(args.steps, 40))

    # Every step copies ITS inputs host -> device and ITS result device -> host; the copies run on two copy streams
    # (one per direction = one per copy engine) so that step i+1's inputs and step i-1's result move under step i's
    # kernels — the way a serving loop drives VideoDecoder.p_sample.  Device buffers are double-buffered; events order a
    # slot's refill after the step that read it, and a result's copy-out after the step that wrote it.
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    slots = [dict(x=torch.empty(shape, device=dev), e=torch.empty_like(emb),
                  ht=torch.empty(B, dtype=torch.long).pin_memory(), hout=torch.empty(shape).pin_memory(),
                  staged=torch.cuda.Event(), used=torch.cuda.Event(), out_done=torch.cuda.Event()) for _ in range(2)]

    def stage_in(i):
        sl = slots[i % 2]
        sl["ht"].copy_(ht - i)
        with torch.cuda.stream(s_in):
            s_in.wait_event(sl["used"])            # the step that last read this slot has finished
            sl["x"].copy_(hx, non_blocking=True)
            sl["e"].copy_(hemb, non_blocking=True)
            sl["staged"].record(s_in)

    def e2e_run(n):
        for sl in slots:
            sl["used"].record(main)
            sl["out_done"].record(s_out)
        stage_in(0)
        outs = [None, None]   # only the two results in flight stay alive (record_stream covers their copies): keeping every
        #                       step's 50 MB result made each p_sample a fresh cudaMalloc, which synchronises the device
        for i in range(n):
            sl = slots[i % 2]
            if i + 1 < n:
                stage_in(i + 1)
            main.wait_event(sl["staged"])
            # t is handed over as the pinned host tensor (p_sample moves it itself): no device read-back inside the call
            out, _ = dec.p_sample(unet, sl["x"], sl["ht"], sl["e"], sched)
            sl["used"].record(main)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                sl["hout"].copy_(out, non_blocking=True)
                out.record_stream(s_out)
                sl["out_done"].record(s_out)
            outs[i % 2] = out
            if i >= 1:
                slots[(i - 1) % 2]["out_done"].synchronize()   # the host consumes step i-1's result while step i runs
        slots[(n - 1) % 2]["out_done"].synchronize()
        torch.cuda.synchronize(dev)

    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    hout = slots[(e2e_steps - 1) % 2]["hout"]
    assert bool(torch.isfinite(hout).all()), "non-finite end-to-end result"

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([elapsed, e2e_elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, e2e_elapsed = float(t[0]), float(t[1])

    conv_profile = profile_conv_family(eng) if rank == 0 else None
    ws_gb = eng.pool.total / 1e9
    # the extra legs need the memory of the sampling plan (and must not disturb it while it is timed)
    del graph, eng, params
    unet._engines.clear()
    torch.cuda.empty_cache()
    train = None
    if not args.no_train:
        train = train_block(dev, world, rank, args.train_batch, max(3, min(args.steps, 10)), 3)
    eager = None
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        eager = gpu_eager_baseline(dev, B)

    if rank == 0:
        pk = peaks()
        flops, conv_sec, n_conv, n_tc, dom_sec, dom_flops, dom_n, (n_fused, fused_sec, fused_bytes) = conv_profile
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get(f"conv_halo_3x3_64x64_C64_C64_B{B}")
        achieved = flops / conv_sec / 1e12
        value = world * B * args.steps / elapsed
        e2e_val = world * B * e2e_steps / e2e_elapsed
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sec, threads, kind, what = cpu_reference_step_time(3, 1)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"3 denoise steps at batch 1 after 1 warm-up: {what}, fp32, torch CPU, all host threads"}
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
                    "api": "VideoDecoder.p_sample(unet, x, t, video_embed, scheduler) per step with pinned host x/t/embed in and "
                           "x_(t-1) out EVERY step; the copies run on two copy streams (double-buffered device inputs), so "
                           "step i+1's H2D and step i-1's D2H overlap step i's kernels"},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor",
                         "kernel": "dv_conv_igemm (the tensor-bound launches of one step: the 3x3 / 1x1 convolutions AND the q / "
                                   "kv / out projections of the attention blocks, which run through the same entry point; the "
                                   "HBM-bound 1x1 res_conv launches that also normalise their residual operand are reported in "
                                   "hbm_fused_residual)",
                         "achieved": achieved,
                         "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                         "peak_burst": pk["bf16_tflops"], "frac_burst": achieved / pk["bf16_tflops"],
                         "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside the step); frac_burst "
                                        "is against the burst figure — MEASURED_PEAKS' sustained number was taken at a lower "
                                        "SM clock than this bench settles at, so both are reported",
                         "algorithmic_gflop_per_step": flops / 1e9, "conv_ms_per_step": conv_sec * 1e3,
                         "launches": n_conv, "tcgen05_launches": n_tc,
                         "dominant_launch": {"kernel": "conv_halo_kernel 3x3 64->64 @64x64 (dv_conv_igemm)", "launches_per_step": dom_n,
                                             "algorithmic_gflop": dom_flops / 1e9, "avg_ms": dom_sec * 1e3,
                                             "achieved": dom_flops / dom_sec / 1e12 if dom_sec else None,
                                             "frac": dom_flops / dom_sec / 1e12 / pk["bf16_sustained"] if dom_sec else None,
                                             "frac_burst": dom_flops / dom_sec / 1e12 / pk["bf16_tflops"] if dom_sec else None,
                                             "algorithmic_bytes": 2.0 * 2 * B * FRAMES * SIZE * SIZE * 64},
                         "hbm_fused_residual": {
                             "kernel": "conv_tc_kernel with add_* (1x1 res_conv + SiLU(GroupNorm(h)) of the residual operand: the "
                                       "tail of every ResnetBlock3D with a res_conv)",
                             "launches_per_step": n_fused, "ms_per_step": fused_sec * 1e3,
                             "algorithmic_gbytes_per_step": fused_bytes / 1e9,
                             "achieved_gbs": fused_bytes / fused_sec / 1e9 if fused_sec else None,
                             "peak_gbs": pk["hbm_gbs"],
                             "frac": fused_bytes / fused_sec / 1e9 / pk["hbm_gbs"] if fused_sec else None},
                         "traffic": traffic,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant launch shape, "
                                         "one ncu --set full capture (profiles/ncu_traffic.json names the summaries); null for batches "
                                         "that were not captured"},
            "cpu_baseline": cpu,
            "gpu_eager_baseline": eager,
            "train": train,
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
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the torch-eager-on-the-same-GPU leg (N=1 only)")
    ap.add_argument("--no-train", action="store_true", help="skip the config-4 (data-parallel training step) block")
    ap.add_argument("--train-batch", type=int, default=32,
                    help="videos per GPU in the training block (measured 1 GPU: 8 -> 330, 16 -> 418, 32 -> 463 samples/s; 56 GB at 32)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
