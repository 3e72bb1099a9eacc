"""CPU, world_size 2 over gloo: batch sharding of the sampler equals the single-process result
(the N>1 path has no data-path collective; the gather here is only the test's comparison)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    import json
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dalle2_video_b200.parallel import shard_rows
    from oracle.unet3d import UnetConfig, unet3d_forward
    from oracle.weights import synth_state_dict
    torch.set_num_threads(2)
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    lay = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_layouts.json")))["tiny"]
    sd = synth_state_dict(lay, seed=7)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 3, 2, 16, 16, generator=g)
    emb = torch.randn(4, 512, generator=g)
    t = torch.tensor([3, 3, 3, 3])
    with torch.no_grad():
        mine = unet3d_forward(sd, UnetConfig(**kw), shard_rows(x, rank, world), shard_rows(t, rank, world),
                              video_embed=shard_rows(emb, rank, world))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    if rank == 0:
        with torch.no_grad():
            full = unet3d_forward(sd, UnetConfig(**kw), x, t, video_embed=emb)
        q.put(float((torch.cat(parts) - full).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_rows_equal_full_batch_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5, err


def _grad_worker(rank, world, port, q):
    import json
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn.functional as F
    from dalle2_video_b200.parallel import allreduce_mean_, shard_rows
    from oracle.unet3d import UnetConfig, unet3d_forward
    from oracle.weights import synth_state_dict
    torch.set_num_threads(2)
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    lay = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_layouts.json")))["tiny"]
    g = torch.Generator().manual_seed(9)
    x, noise = torch.randn(4, 3, 2, 16, 16, generator=g), torch.randn(4, 3, 2, 16, 16, generator=g)
    emb, t = torch.randn(4, 512, generator=g), torch.tensor([3, 40, 7, 99])

    def flat_grad(xs, ns, es, ts):
        sd = {k: v.requires_grad_() for k, v in synth_state_dict(lay, seed=7).items()}
        pred = unet3d_forward(sd, UnetConfig(**kw), xs, ts, video_embed=es)
        F.mse_loss(pred, ns, reduction="none").flatten(1).mean(1).mean().backward()
        return torch.cat([sd[k].grad.reshape(-1) if sd[k].grad is not None else torch.zeros(sd[k].numel()) for k in sorted(sd)])

    mine = flat_grad(*(shard_rows(v, rank, world) for v in (x, noise, emb, t)))
    allreduce_mean_(mine)   # the trainer's one collective: equal shards => mean of shard gradients = full-batch gradient
    if rank == 0:
        full = flat_grad(x, noise, emb, t)
        q.put(float((mine - full).norm() / full.norm()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_equals_full_batch_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5, err


def test_grad_buckets_partition_and_merge():
    """Bucket construction (no process group): the buckets tile the flat buffer exactly, each is ready no earlier than
    every parameter inside it, small runs are merged, and parameters nobody claims wait for finish()."""
    from dalle2_video_b200.parallel import GradBuckets
    sizes = [300, 5, 5000, 7000, 3, 2, 9000, 40, 6000]
    segs = [4, 9, 3, 3, 2, 2, 1, 0, 0]       # backward finishes the tail of the buffer first; item 1 arrives via autograd
    items, off = [], 0
    for n, s in zip(sizes, segs):
        items.append((off, n, s))
        off += n
    flat = torch.zeros(off + 17)             # 17 trailing elements no parameter owns
    gb = GradBuckets(flat, items, n_segments=5, min_bytes=4 * 1000)
    assert gb.buckets[0][0] == 0 and gb.buckets[-1][1] == flat.numel()
    assert all(a[1] == b[0] for a, b in zip(gb.buckets, gb.buckets[1:]))
    for lo, hi, seg in gb.buckets:
        inside = [min(s, 5) for (o, n, s) in items if o < hi and o + n > lo]
        assert seg >= max(inside or [5])
        assert (hi - lo) * 4 >= 4000 or seg == 5   # only buckets that wait for finish() anyway may stay small
    assert len(gb.buckets) <= 5


def test_small_runs_of_the_last_segment_do_not_delay_their_neighbours():
    """The conditioning prologue's parameters (every time_mlp / to_kv weight) become final with the LAST backward segment and
    sit between early-finished convolution weights in the flat buffer: merged into a neighbour they would push most of the
    gradient behind the end of the backward; they stay buckets of their own, as late (autograd) arrivals always did."""
    from dalle2_video_b200.parallel import GradBuckets
    sizes = [6000, 50, 7000, 30, 5000]
    segs = [0, 2, 1, 2, 0]
    items, off = [], 0
    for n, s in zip(sizes, segs):
        items.append((off, n, s))
        off += n
    gb = GradBuckets(torch.zeros(off), items, n_segments=3, min_bytes=4 * 1000)
    assert [b[2] for b in gb.buckets] == [0, 2, 1, 2, 0]
    early = sum(hi - lo for lo, hi, seg in gb.buckets if seg < 2)
    assert early == 18000


def _bucket_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dalle2_video_b200.parallel import GradBuckets, allreduce_mean_
    g = torch.Generator().manual_seed(100 + rank)
    sizes, segs = [4096, 64, 8192, 2048, 16, 4096], [3, 3, 2, 1, 0, 0]
    items, off = [], 0
    for n, s in zip(sizes, segs):
        items.append((off, n, s))
        off += n
    flat = torch.zeros(off)
    gb = GradBuckets(flat, items, n_segments=3, min_bytes=1024)
    final = torch.randn(off, generator=g)
    want = allreduce_mean_(final.clone())
    for seg in range(3):                      # "backward": segment `seg` fills in the slices it owns, then launches
        for o, n, s in items:
            if s == seg:
                flat[o:o + n] = final[o:o + n]
        gb.launch(seg)
    for o, n, s in items:                     # late arrivals (the prologue parameters of the trainer)
        if s == 3:
            flat[o:o + n] = final[o:o + n]
    gb.finish()
    first = float((flat - want).abs().max())
    flat.copy_(final)                         # second step on the same object: state was reset by finish()
    for seg in range(3):
        gb.launch(seg)
    gb.finish()
    if rank == 0:
        q.put(max(first, float((flat - want).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_equals_one_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6, err


def _tiny_trainer(seed, **kw):
    """A VideoDecoderTrainer on CPU tensors (construction touches no kernel) whose decoder forward is replaced by a
    differentiable stand-in, so the trainer's HOST logic can be exercised without a GPU."""
    from dalle2_video_b200 import Unet3D, VideoDecoder, VideoDecoderTrainer
    torch.manual_seed(seed)
    unet = Unet3D(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                  cond_on_text_encodings=True, text_embed_dim=64)
    dec = VideoDecoder(unet, frame_size=16, frame_numbers=(2,), timesteps=10, learned_variance=False)
    tr = VideoDecoderTrainer(dec, lr=1e-3, **kw)
    w = dec.unets[0].to_time_cond[0].weight

    def fake_forward(video, video_embed=None, unet_number=None, **_):
        return (w[: video.shape[1]].sum(1)[None, :, None] * video).pow(2).mean()

    dec.forward = fake_forward
    return tr, dec, w


def _sync_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    tr, dec, _ = _tiny_trainer(seed=100 + rank)       # replicas built under DIFFERENT seeds
    flat = tr.flat_p.clone()
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    if rank == 0:
        torch.manual_seed(100)
        from dalle2_video_b200 import Unet3D
        ref = Unet3D(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                     cond_on_text_encodings=True, text_embed_dim=64)
        same_as_rank0 = all(torch.equal(p, q_) for p, q_ in zip(ref.parameters(), dec.unets[0].parameters()))
        q.put((float((parts[0] - parts[1]).abs().max()), same_as_rank0))
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_broadcasts_rank0_weights_gloo_world2():
    """accelerate.prepare / DDP broadcast rank 0's parameters at construction (trainer.py:117-124); ranks seeded
    differently must end up with rank 0's weights."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    diff, same = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert diff == 0.0 and same


def test_trainer_eval_forward_leaves_gradients_untouched_and_chunks_accumulate():
    """trainer.py:322-365: backward only `if self.training`; max_batch_size chunks scale the loss by their share, so
    the accumulated gradient equals the one-piece gradient; zero_grad(set_to_none=True) does not detach the views."""
    tr, dec, w = _tiny_trainer(seed=3)
    video = torch.randn(6, 4, 3)
    tr.eval()
    l_eval = tr(video, video_embed=None)
    assert float(tr.flat_g.abs().max()) == 0.0
    tr.train()
    with torch.no_grad():
        assert abs(tr(video, video_embed=None) - l_eval) < 1e-6
    assert float(tr.flat_g.abs().max()) == 0.0
    l_full = tr(video, video_embed=None)
    g_full = tr.flat_g.clone()
    assert abs(l_full - l_eval) < 1e-6 and float(g_full.abs().max()) > 0
    dec.zero_grad()                                   # set_to_none=True: detaches param.grad from the flat buffer
    assert w.grad is None
    l_chunk = tr(video, video_embed=None, max_batch_size=4)   # chunks of 4 + 2
    assert w.grad is not None and w.grad.data_ptr() >= tr.flat_g.data_ptr()
    assert abs(l_chunk - l_full) < 1e-5
    # mean over a chunk x share of the batch = mean over the batch (the stand-in loss is a plain mean over samples)
    assert float((tr.flat_g - g_full).abs().max()) < 1e-5 * float(g_full.abs().max()) + 1e-8


def test_trainer_rejects_or_names_unsupported_keywords():
    import pytest
    import warnings
    from dalle2_video_b200 import Unet3D, VideoDecoder, VideoDecoderTrainer
    unet = Unet3D(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    dec = VideoDecoder(unet, frame_size=16, frame_numbers=(2,), timesteps=10, learned_variance=False)
    with pytest.raises(NotImplementedError):
        VideoDecoderTrainer(dec, use_ema=True)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        tr = VideoDecoderTrainer(dec, ema_beta=0.99, warmup_steps=10, cosine_decay_max_steps=100)
    assert any("ema_beta" in str(r.message) for r in rec)
    # LinearWarmup dampening x CosineAnnealingLR (trainer.py:78-90): first step runs at lr/10
    assert abs(tr.current_lr() - 1e-4 * 0.1) < 1e-12
    tr.steps[0] = 50
    assert abs(tr.current_lr() - 1e-4 * 0.5) < 1e-12


def test_adamw_ranges_skip_parameters_no_training_path_reaches():
    """torch.optim.AdamW skips parameters whose grad is None (no weight decay either): null_text_embed / text_to_cond
    never get a gradient on this path, so they sit outside the ranges dv_adamw steps."""
    tr, dec, _ = _tiny_trainer(seed=5)
    st = tr._opt[0]
    stepped = torch.zeros(st.flat_p.numel(), dtype=torch.bool)
    for off, n, _ in st.groups:
        stepped[off:off + n] = True
    base = st.flat_p.data_ptr()
    for name, p in dec.unets[0].named_parameters():
        off = (p.data_ptr() - base) // 4
        want = not name.startswith(("null_text_embed", "text_to_cond."))
        assert bool(stepped[off:off + p.numel()].all()) == want and bool(stepped[off:off + p.numel()].any()) == want, name
    # weight decay only on the >= 2-d parameters (get_optimizer's group_wd_params)
    (o0, n0, wd0), (o1, n1, wd1) = st.groups
    assert wd0 > 0 and wd1 == 0.0
