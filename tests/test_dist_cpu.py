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
