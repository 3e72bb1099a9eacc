"""-m gpu: the decoder training step (BASELINE.json configs[3]) against torch autograd through
the CPU oracle (oracle/unet3d.py) on the same weights, noise, timesteps and keep-masks.  Activations are bf16 end to end,
so the bar is statistical: loss within 2 %, every large gradient tensor with cosine similarity >= 0.995 to the fp32
oracle gradient, every tensor's norm within 3 % of the oracle's, the whole gradient within 2 % in relative L2 norm
(measured 5.5e-3; the round-1 bounds of 0.98 / 8 % would have passed a mis-scaled tensor)."""
import pytest
import torch
import torch.nn.functional as F

from dalle2_video_b200 import Unet3D, VideoDecoder, VideoDecoderTrainer
from dalle2_video_b200.train import p_losses
from gpu_util import rel_err
from oracle.primitives import NoiseScheduler as OracleScheduler
from oracle.unet3d import UnetConfig, unet3d_forward
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu

CASES = {
    "tiny": (dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True), (2, 3, 4, 16, 16)),
    "headline_small": (dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True), (2, 3, 4, 16, 16)),
    # memory_efficient (:466-541, :611): initial resnet block, downsample before each stage's blocks, upsample after every up stage
    "memory_efficient": (dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True, memory_efficient=True),
                         (2, 3, 4, 32, 32)),
}


def build(kw, frames, size, seed=5):
    torch.manual_seed(0)
    u = Unet3D(**kw)
    layout = {k: list(v.shape) for k, v in u.state_dict().items()}
    sd = synth_state_dict(layout, seed=seed)
    u.load_state_dict(sd, strict=True)
    dec = VideoDecoder(u, frame_size=size, frame_numbers=(frames,), timesteps=100, learned_variance=False,
                       video_cond_drop_prob=0.0, text_cond_drop_prob=0.0).cuda()
    return dec, sd


def oracle_loss_and_grads(sd, kw, video, times, noise, emb, drop_video=False):
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    sched = OracleScheduler(beta_schedule="cosine", timesteps=100, loss_type="l2")
    x_t = sched.q_sample(2 * video - 1, times, noise)
    pred = unet3d_forward(sdg, UnetConfig(**kw), x_t, times, video_embed=emb, drop_video=drop_video)
    loss = F.mse_loss(pred, noise, reduction="none").flatten(1).mean(1).mean()
    loss.backward()
    return float(loss), {k: v.grad for k, v in sdg.items() if v.grad is not None}


@pytest.mark.parametrize("case", list(CASES))
def test_training_loss_and_gradients_match_oracle_autograd(case):
    kw, shape = CASES[case]
    dec, sd = build(kw, shape[2], shape[3])
    g = torch.Generator().manual_seed(3)
    video = torch.rand(shape, generator=g)
    noise = torch.randn(shape, generator=g)
    emb = torch.randn(shape[0], 512, generator=g)
    times = torch.tensor([7, 63])
    ref_loss, ref_grads = oracle_loss_and_grads(sd, kw, video, times, noise, emb)
    unet = dec.unets[0]
    loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0],
                    noise=noise.cuda())
    assert abs(float(loss) - ref_loss) < 2e-2 * ref_loss, (float(loss), ref_loss)
    loss.backward()
    got = {k: p.grad for k, p in unet.named_parameters()}
    missing = [k for k in ref_grads if got.get(k) is None]
    assert not missing, missing
    num = den = 0.0
    worst = (1.0, None)
    total = sum(float(rg.pow(2).sum()) for rg in ref_grads.values()) ** 0.5
    ratios = []
    for k, rg in ref_grads.items():
        gg = got[k].float().cpu()
        assert gg.shape == rg.shape, k
        num += float((gg - rg).pow(2).sum())
        den += float(rg.pow(2).sum())
        if rg.numel() >= 256 and float(rg.norm()) > 1e-6:
            cos = float(F.cosine_similarity(gg.flatten(), rg.flatten(), dim=0))
            if cos < worst[0]:
                worst = (cos, k)
        if float(rg.norm()) > 1e-3 * total:   # every tensor that carries a visible share of the gradient
            ratios.append((float(gg.norm() / rg.norm()), k))
    # measured (round 2): total rel-L2 5.5e-3 / 5.3e-3, worst cosine 0.9996, norm ratios 0.992 .. 1.005
    assert worst[0] >= 0.995, worst
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5
    # a mis-scaled tensor (a dropped factor in one backward kernel) keeps its cosine: bound every tensor's norm as well
    lo, hi = min(ratios), max(ratios)
    print(f"{case}: total rel-L2 {(num / den) ** 0.5:.3e}, worst cosine {worst}, norm ratio range {lo} .. {hi} over {len(ratios)} tensors")
    assert 0.97 < lo[0] and hi[0] < 1.03, (lo, hi)


@pytest.mark.parametrize("case", list(CASES))
def test_training_step_in_float32_parity_mode_matches_oracle_autograd(case):
    """unet.compute_dtype = torch.float32: the training plan in float32 on the CUDA-core kernels (activations, weight copies,
    activation gradients, the CUDA-core weight-gradient kernel) — north_star's "fp32 within 1e-4 per step" on the training
    side (dalle2_video.py:1908-2006): loss within 1e-5 relative, every gradient tensor within 1e-4 relative L2 of autograd
    through the CPU oracle, the whole gradient within 1e-5 (measured on B200: 5.8e-7 / 7.5e-7 for the whole gradient, worst
    tensor 4.9e-6, loss equal to 8 digits)."""
    kw, shape = CASES[case]
    dec, sd = build(kw, shape[2], shape[3])
    g = torch.Generator().manual_seed(3)
    video = torch.rand(shape, generator=g)
    noise = torch.randn(shape, generator=g)
    emb = torch.randn(shape[0], 512, generator=g)
    times = torch.tensor([7, 63])
    ref_loss, ref_grads = oracle_loss_and_grads(sd, kw, video, times, noise, emb)
    unet = dec.unets[0]
    unet.compute_dtype = torch.float32
    loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0],
                    noise=noise.cuda())
    assert abs(float(loss) - ref_loss) < 1e-5 * ref_loss, (float(loss), ref_loss)
    loss.backward()
    got = {k: p.grad for k, p in unet.named_parameters()}
    num = den = 0.0
    total = sum(float(rg.pow(2).sum()) for rg in ref_grads.values()) ** 0.5
    worst = (0.0, None)
    for k, rg in ref_grads.items():
        gg = got[k].float().cpu()
        num += float((gg - rg).pow(2).sum())
        den += float(rg.pow(2).sum())
        if float(rg.norm()) > 1e-4 * total:
            e = float((gg - rg).norm() / rg.norm())
            if e > worst[0]:
                worst = (e, k)
    print(f"{case} fp32: loss {float(loss):.7f} vs {ref_loss:.7f}, total rel-L2 {(num / den) ** 0.5:.3e}, worst tensor {worst}")
    assert (num / den) ** 0.5 < 1e-5, (num / den) ** 0.5
    assert worst[0] < 1e-4, worst


def test_dropped_video_conditioning_routes_gradients_to_the_null_embeddings():
    kw, shape = CASES["tiny"]
    dec, sd = build(kw, shape[2], shape[3])
    dec.video_cond_drop_prob = 1.0   # keep-mask all False, no RNG consumed (prob_mask_like)
    g = torch.Generator().manual_seed(4)
    video, noise = torch.rand(shape, generator=g), torch.randn(shape, generator=g)
    emb, times = torch.randn(shape[0], 512, generator=g), torch.tensor([1, 99])
    ref_loss, ref_grads = oracle_loss_and_grads(sd, kw, video, times, noise, emb, drop_video=True)
    unet = dec.unets[0]
    loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0],
                    noise=noise.cuda())
    loss.backward()
    assert abs(float(loss) - ref_loss) < 2e-2 * ref_loss
    for k in ("null_video_embed", "null_video_hiddens"):
        gk = dict(unet.named_parameters())[k].grad
        assert rel_err(gk, ref_grads[k]) < 8e-2, k
    assert float(unet.video_to_tokens[0].weight.grad.abs().max()) == 0.0


def test_trainer_steps_reduce_the_loss_and_follow_the_reference_rng_order():
    kw, shape = CASES["tiny"]
    dec, _ = build(kw, shape[2], shape[3])
    dec.video_cond_drop_prob, dec.text_cond_drop_prob = 0.1, 0.5
    trainer = VideoDecoderTrainer(dec, lr=3e-4, wd=1e-2, max_grad_norm=0.5)
    g = torch.Generator().manual_seed(9)
    video, emb = torch.rand(shape, generator=g).cuda(), torch.randn(shape[0], 512, generator=g).cuda()
    # RNG contract: randint(times) -> randn_like(noise) -> uniform (video mask) -> uniform (text mask)
    torch.manual_seed(123)
    l0 = trainer(video, video_embed=emb)
    after = torch.rand(1, device="cuda")
    torch.manual_seed(123)
    torch.randint(0, 100, (shape[0],), device="cuda")
    torch.randn_like(video)
    torch.zeros(shape[0], device="cuda").uniform_(0, 1)
    torch.zeros(shape[0], device="cuda").uniform_(0, 1)
    assert torch.equal(after, torch.rand(1, device="cuda"))
    trainer.update()
    losses = [l0]
    for _ in range(12):
        torch.manual_seed(123)   # same timesteps / noise / masks: the loss on this batch must go down
        losses.append(trainer(video, video_embed=emb))
        trainer.update()
    assert int(trainer.steps[0]) == 13
    assert losses[-1] < 0.9 * losses[0], losses
    assert all(torch.isfinite(p).all() for p in dec.parameters())
    # sampling still works on the updated weights (plans are rebuilt)
    out = dec.sample(video_embed=emb)
    assert out.shape == video.shape and bool(torch.isfinite(out).all())


def test_split_backward_gives_the_same_gradients_and_fires_segments_in_order():
    """The data-parallel trainer cuts the backward into segments (one CUDA graph each) and starts a bucket's all-reduce
    after the segment that finishes it.  Here, on one GPU: the segmented backward reproduces the
    one-piece gradients on the eager, capture and replay calls, every segment reports exactly once and in
    order, and the buckets built from the plan tile the flat buffer with ready-segments that respect every parameter."""
    from dalle2_video_b200.parallel import GradBuckets
    kw, shape = CASES["headline_small"]
    g = torch.Generator().manual_seed(3)
    video, noise = torch.rand(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
    emb, times = torch.randn(shape[0], 512, generator=g).cuda(), torch.tensor([7, 63]).cuda()

    def grads(split):
        dec, _ = build(kw, shape[2], shape[3])
        unet = dec.unets[0]
        fired, out = [], []
        for _ in range(3):   # eager, capture, replay
            for p in unet.parameters():
                p.grad = None
            plan = unet.train_plan(shape[0], shape[2], shape[3], shape[4], video.device)
            if split and plan.segments is None:
                nseg = plan.split_backward(bucket_bytes=8 << 20, tail_bytes=1 << 20)
                assert nseg >= 4, nseg
                plan.on_segment = fired.append
            loss = p_losses(dec, unet, video, times, video_embed=emb, noise_scheduler=dec.noise_schedulers[0], noise=noise)
            loss.backward()
            out.append({k: p.grad.clone() for k, p in unet.named_parameters() if p.grad is not None})
        return plan, fired, out

    _, _, whole = grads(False)
    plan, fired, parts = grads(True)
    nseg = len(plan.segments)
    assert fired == list(range(nseg)) * 3
    assert plan.segments[0][0] == 0 and plan.segments[-1][1] == len(plan.bwd)
    assert all(a[1] == b[0] for a, b in zip(plan.segments, plan.segments[1:]))
    for a, b in zip(whole, parts):
        assert a.keys() == b.keys()
        # same launches in the same order; the fp32 atomics of the attention / norm reductions reorder from run to run
        # and the bf16 activations downstream amplify that to a few 1e-3 (a few 1e-2 on the smallest gradients) — a
        # dropped or misplaced segment would be an O(1) error in the tensors it owns
        num = sum(float((a[k] - b[k]).pow(2).sum()) for k in a)
        den = sum(float(a[k].pow(2).sum()) for k in a)
        assert (num / den) ** 0.5 < 1e-2, (num / den) ** 0.5
        for k in a:
            if float(a[k].norm()) > 1e-7:
                cos = float(F.cosine_similarity(a[k].flatten(), b[k].flatten(), dim=0))
                assert cos > 0.995, (k, cos)
    # bucket map over a flat buffer laid out like the trainer's
    prms = [p for p in plan.u.parameters()]
    seg_of = {}
    for s, (_, _, ids) in enumerate(plan.segments):
        for i in ids:
            seg_of[id(plan.grad_map[i][0])] = max(seg_of.get(id(plan.grad_map[i][0]), 0), s)
    items, off = [], 0
    for p in prms:
        items.append((off, p.numel(), seg_of.get(id(p), nseg)))
        off += p.numel()
    gb = GradBuckets(torch.zeros(off), items, nseg, min_bytes=1 << 20)
    assert gb.buckets[0][0] == 0 and gb.buckets[-1][1] == off
    for lo, hi, seg in gb.buckets:
        assert seg >= max(s for (o, n, s) in items if o < hi and o + n > lo)
    early = sum(hi - lo for lo, hi, seg in gb.buckets if seg < nseg - 1)
    assert early > 0.8 * off, (early, off)   # most of the gradient travels before the last segment has run


@pytest.mark.parametrize("name", ["v", "x_start"])
def test_training_loss_of_x_start_and_v_objectives_matches_reference_fixture(name, layouts):
    """p_losses target selection (:1990-1995) — dv_q_sample writes x_start / v next to x_t; loss against the reference's
    own p_losses on the same video, noise and timesteps (bf16 activations: 2 %)."""
    from conftest import golden
    g = golden("sample_objectives.pt")
    case = g["cases"][name]
    kw, _ = CASES["tiny"]
    u = Unet3D(**kw)
    u.load_state_dict(synth_state_dict(layouts["tiny"], seed=g["weight_seed"]), strict=True)
    dec = VideoDecoder(u, frame_size=16, frame_numbers=(4,), timesteps=g["timesteps"], learned_variance=False,
                       video_cond_drop_prob=0.0, text_cond_drop_prob=0.0, **case["decoder_kwargs"]).cuda()
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(g["video_embed_seed"]))
    gen = torch.Generator().manual_seed(g["loss_input_seed"])
    video, noise = torch.rand(g["shape"], generator=gen), torch.randn(g["shape"], generator=gen)
    times = torch.tensor(g["loss_times"])
    loss = p_losses(dec, dec.unets[0], video.cuda(), times.cuda(), video_embed=emb.cuda(),
                    noise_scheduler=dec.noise_schedulers[0], noise=noise.cuda(),
                    predict_x_start=dec.predict_x_start[0], predict_v=dec.predict_v[0])
    assert abs(float(loss) - case["loss"]) < 2e-2 * case["loss"], (float(loss), case["loss"])
    loss.backward()
    grads = [p.grad for p in dec.unets[0].parameters() if p.grad is not None]
    assert len(grads) > 100 and all(bool(torch.isfinite(gr).all()) for gr in grads)
    assert float(dec.unets[0].to_out.weight.grad.abs().max()) > 0


@pytest.mark.parametrize("label", ["with_first_pass", "without_first_pass"])
def test_self_conditioned_training_loss_matches_reference_fixture(label, layouts):
    """p_losses with Unet3D(self_cond=True) (:1969-1983): `random.random() < 0.5` decides whether a no-grad first pass
    supplies the self-conditioning planes; both branches against the reference's loss (bf16: 2 %)."""
    import random
    from conftest import golden
    g = golden("sample_objectives.pt")
    case = g["cases"]["self_cond"]
    u = Unet3D(**case["unet_kwargs"])
    u.load_state_dict(synth_state_dict(layouts["tiny_self_cond"], seed=g["weight_seed"]), strict=True)
    dec = VideoDecoder(u, frame_size=16, frame_numbers=(4,), timesteps=g["timesteps"], learned_variance=False,
                       video_cond_drop_prob=0.0, text_cond_drop_prob=0.0).cuda()
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(g["video_embed_seed"]))
    gen = torch.Generator().manual_seed(g["loss_input_seed"])
    video, noise = torch.rand(g["shape"], generator=gen), torch.randn(g["shape"], generator=gen)
    random.seed(case["python_random_seeds"][label])
    loss = p_losses(dec, dec.unets[0], video.cuda(), torch.tensor(g["loss_times"]).cuda(), video_embed=emb.cuda(),
                    noise_scheduler=dec.noise_schedulers[0], noise=noise.cuda())
    want = case["loss"][label]
    assert abs(float(loss) - want) < 2e-2 * want, (float(loss), want)
    loss.backward()
    gw = dec.unets[0].init_conv.convs[0].weight.grad
    assert gw is not None and gw.shape[1] == 6 and float(gw[:, 3:].abs().max()) > 0 or label == "without_first_pass"


@pytest.mark.parametrize("kind", ["l2", "l1", "huber"])
@pytest.mark.parametrize("weighted", [False, True])
def test_loss_kernel_kinds_and_p2_weights_match_torch_autograd(kind, weighted):
    """dv_mse_loss = noise_scheduler.loss_fn(pred, target, 'none') -> per-sample mean -> p2_reweigh_loss -> mean (:1997-2003)."""
    from dalle2_video_b200 import _lib as L
    from gpu_util import ptr, stream
    g = torch.Generator().manual_seed(5)
    pred = (torch.randn(3, 3, 4, 8, 8, generator=g) * 1.5).cuda().requires_grad_()
    target = torch.randn(3, 3, 4, 8, 8, generator=g).cuda()
    w = torch.tensor([0.3, 1.0, 2.5]).cuda() if weighted else None
    fn = {"l2": F.mse_loss, "l1": F.l1_loss, "huber": F.smooth_l1_loss}[kind]
    per = fn(pred, target, reduction="none").flatten(1).mean(1)
    ref = (per * w if weighted else per).mean()
    ref.backward()
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    dpred = torch.empty_like(target)
    L.call("dv_mse_loss", L.MseLossParams(pred=ptr(pred), target=ptr(target), dpred=ptr(dpred), loss=ptr(loss), n=pred.numel(),
                                          grad_scale=0.7, kind=L.LOSS_KINDS[kind], sample_weight=ptr(w) if weighted else None,
                                          per_sample=pred.numel() // 3), stream())
    assert abs(float(loss) - float(ref)) < 1e-5 * float(ref)
    assert rel_err(dpred, 0.7 * pred.grad) < 1e-5


def test_huber_loss_with_p2_reweighting_through_the_decoder(layouts):
    kw, _ = CASES["tiny"]
    u = Unet3D(**kw)
    sd = synth_state_dict(layouts["tiny"], seed=7)
    u.load_state_dict(sd, strict=True)
    dec = VideoDecoder(u, frame_size=16, frame_numbers=(4,), timesteps=100, learned_variance=False, loss_type="huber",
                       p2_loss_weight_gamma=1.0, video_cond_drop_prob=0.0, text_cond_drop_prob=0.0).cuda()
    g = torch.Generator().manual_seed(3)
    shape = (2, 3, 4, 16, 16)
    video, noise, emb = torch.rand(shape, generator=g), torch.randn(shape, generator=g), torch.randn(2, 512, generator=g)
    times = torch.tensor([7, 63])
    sched = OracleScheduler(beta_schedule="cosine", timesteps=100, loss_type="huber", p2_loss_weight_gamma=1.0)
    with torch.no_grad():
        pred = unet3d_forward(sd, UnetConfig(**kw), sched.q_sample(2 * video - 1, times, noise), times, video_embed=emb)
    per = F.smooth_l1_loss(pred, noise, reduction="none").flatten(1).mean(1)
    ref = float(sched.p2_reweigh_loss(per, times).mean())
    loss = p_losses(dec, dec.unets[0], video.cuda(), times.cuda(), video_embed=emb.cuda(),
                    noise_scheduler=dec.noise_schedulers[0], noise=noise.cuda())
    assert abs(float(loss) - ref) < 2e-2 * ref, (float(loss), ref)
    loss.backward()
    assert float(dec.unets[0].to_out.weight.grad.abs().max()) > 0


@pytest.mark.parametrize("key", ["cascade_train_u2", "cascade_train_u2_noise"])
def test_cascade_stage2_training_matches_reference_loss_and_oracle_gradients(key, layouts):
    """Training the second unet of a cascade (scope rows a16 x N2): lowres planes into the init conv (:744-746) and the head
    (:949-952, dv_to_out_bwd `extra`), noise-level conditioning into t (:765-768).  Loss against the reference's own
    p_losses, gradients against torch autograd through the oracle."""
    from conftest import golden
    from dalle2_video_b200.decoder import LowresVideoConditioner
    g = golden("cascade_train.pt")
    case = g["cases"][key]
    kw = case["unet_kwargs"]
    use_noise = case["lowres_noise_level"] is not None
    sd = synth_state_dict(layouts[key], seed=g["weight_seed"])
    u1, u2 = Unet3D(**CASES["tiny"][0]), Unet3D(**kw)
    dec = VideoDecoder((u1, u2), frame_sizes=(16, 32), frame_numbers=(4, 4), timesteps=g["timesteps"], learned_variance=False,
                       use_noise_for_lowres_cond=use_noise, video_cond_drop_prob=0.0, text_cond_drop_prob=0.0).cuda()
    dec.unets[1].load_state_dict(sd, strict=True)   # after construction: the decoder re-instantiates its unets (quirk Q1)
    gen = torch.Generator().manual_seed(g["input_seed"])
    video, noise = torch.rand(g["shape"], generator=gen), torch.randn(g["shape"], generator=gen)
    emb, times = torch.randn(2, 512, generator=gen), torch.tensor(g["times"])
    cond = LowresVideoConditioner(downsample_first=True, use_blur=False, use_noise=False, input_video_range=(0.0, 1.0))
    low, _ = cond(video.cuda(), target_frame_size=32, downsample_frame_size=16)
    assert torch.equal(low.cpu(), case["lowres_cond_video"])
    level = case["lowres_noise_level"]
    loss = p_losses(dec, dec.unets[1], video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[1],
                    noise=noise.cuda(), lowres_cond_video=low, lowres_noise_level=level.cuda() if use_noise else None)
    assert abs(float(loss) - case["loss"]) < 2e-2 * case["loss"], (float(loss), case["loss"])
    loss.backward()
    # oracle autograd on the same inputs
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    sched = OracleScheduler(beta_schedule="linear", timesteps=g["timesteps"], loss_type="l2")   # last unet: linear (:1376-1385)
    pred = unet3d_forward(sdg, UnetConfig(**kw), sched.q_sample(2 * video - 1, times, noise), times, video_embed=emb,
                          lowres_cond_video=case["lowres_cond_video"] * 2 - 1, lowres_noise_level=level)
    F.mse_loss(pred, noise, reduction="none").flatten(1).mean(1).mean().backward()
    got = dict(dec.unets[1].named_parameters())
    num = den = 0.0
    for k, v in sdg.items():
        if v.grad is None:
            continue
        gg = got[k].grad.float().cpu()
        num += float((gg - v.grad).pow(2).sum())
        den += float(v.grad.pow(2).sum())
        if v.grad.numel() >= 256 and float(v.grad.norm()) > 1e-6:
            cos = float(F.cosine_similarity(gg.flatten(), v.grad.flatten(), dim=0))
            assert cos >= 0.98, (k, cos)
    assert (num / den) ** 0.5 < 8e-2
    for k in ["to_out.weight"] + (["to_lowres_noise_cond.3.weight"] if use_noise else []):
        assert rel_err(got[k].grad, sdg[k].grad) < 8e-2, k
    wg, rg = got["to_out.weight"].grad.float().cpu(), sdg["to_out.weight"].grad
    assert rel_err(wg[:, -3:], rg[:, -3:]) < 5e-2    # the columns fed by the low-resolution planes
    ig, irg = got["init_conv.convs.0.weight"].grad.float().cpu(), sdg["init_conv.convs.0.weight"].grad
    assert ig.shape[1] == 6 and rel_err(ig[:, 3:], irg[:, 3:]) < 8e-2


def test_cascade_stage2_trainer_steps_through_decoder_forward(layouts):
    """VideoDecoder.forward(unet_number=2) builds the conditioning video itself (nearest down / up + noise augmentation at a
    random level, :2245-2258) from a video larger than the stage's frame size; a few optimiser steps reduce the loss."""
    kw = dict(CASES["tiny"][0], lowres_cond=True)
    u1, u2 = Unet3D(**CASES["tiny"][0]), Unet3D(**kw)
    dec = VideoDecoder((u1, u2), frame_sizes=(16, 32), frame_numbers=(4, 4), timesteps=100, learned_variance=False,
                       use_noise_for_lowres_cond=True, video_cond_drop_prob=0.0, text_cond_drop_prob=0.0).cuda()
    torch.nn.init.normal_(dec.unets[1].to_out.weight, std=0.02)
    g = torch.Generator().manual_seed(2)
    video, emb = torch.rand(2, 3, 4, 64, 64, generator=g).cuda(), torch.randn(2, 512, generator=g).cuda()
    loss, low = dec(video, video_embed=emb, unet_number=2, return_lowres_cond_video=True)
    assert low.shape == (2, 3, 4, 32, 32) and bool(torch.isfinite(loss))
    loss.backward()
    params = [p for p in dec.unets[1].parameters() if p.grad is not None]
    opt = torch.optim.AdamW(params, lr=3e-4)
    losses = []
    for _ in range(10):
        for p in params:
            p.grad = None
        torch.manual_seed(11)
        loss = dec(video, video_embed=emb, unet_number=2)
        loss.backward()
        opt.step()
        dec.unets[1].invalidate_cache()
        losses.append(float(loss))
    assert losses[-1] < 0.9 * losses[0], losses


@pytest.mark.parametrize("k,sigma,shape", [(3, 0.6, (2, 3, 4, 16, 16)), (5, 1.3, (1, 3, 2, 9, 20)), (1, 0.6, (1, 3, 1, 8, 8))])
def test_gaussian_blur_kernel_matches_oracle_restatement(k, sigma, shape):
    from dalle2_video_b200.decoder import LowresVideoConditioner
    from oracle.unet3d import gaussian_blur_frames
    x = torch.rand(shape, generator=torch.Generator().manual_seed(k))
    got = LowresVideoConditioner.blur_image(x.cuda(), sigma, k)
    assert rel_err(got, gaussian_blur_frames(x, k, sigma)) < 1e-6
    if k == 1:
        assert torch.equal(got.cpu(), x)


def test_lowres_conditioner_blurs_half_of_the_time_like_the_reference():
    import random
    from dalle2_video_b200.decoder import LowresVideoConditioner
    from oracle.unet3d import gaussian_blur_frames, resize_nearest
    x = torch.rand(2, 3, 2, 32, 32, generator=torch.Generator().manual_seed(0))
    cond = LowresVideoConditioner(downsample_first=True, use_blur=True, use_noise=False, input_video_range=(0.0, 1.0))
    for seed, blurred in ((1, True), (2, False)):   # random.random() = 0.134 / 0.956 against blur_prob 0.5 (:1138)
        random.seed(seed)
        got, levels = cond(x.cuda(), target_frame_size=32, downsample_frame_size=16)
        low = resize_nearest(x, 16, clamp_range=(0.0, 1.0))
        if blurred:
            low = gaussian_blur_frames(low, 3, 0.6)
        want = resize_nearest(low, 32, clamp_range=(0.0, 1.0))
        assert levels is None and rel_err(got, want) < 1e-6, seed


def test_trainer_keeps_one_optimiser_state_per_unet(layouts):
    """VideoDecoderTrainer.forward / update take the unet number (reference trainer.py:240-274, :322-365): training the
    second unet of a cascade leaves the first one bit-identical (no weight decay, no Adam step on it), each unet has its
    own step counter, and the unet must be named when there are several."""
    kw = CASES["tiny"][0]
    u1, u2 = Unet3D(**kw), Unet3D(**dict(kw, lowres_cond=True))
    dec = VideoDecoder((u1, u2), frame_sizes=(16, 32), frame_numbers=(4, 4), timesteps=100, learned_variance=False,
                       use_noise_for_lowres_cond=True, use_blur_for_lowres_cond=False).cuda()
    torch.nn.init.normal_(dec.unets[1].to_out.weight, std=0.02)
    torch.nn.init.normal_(dec.unets[0].to_out.weight, std=0.02)
    trainer = VideoDecoderTrainer(dec, lr=3e-4, wd=1e-2, max_grad_norm=0.5)
    g = torch.Generator().manual_seed(2)
    video, emb = torch.rand(2, 3, 4, 32, 32, generator=g).cuda(), torch.randn(2, 512, generator=g).cuda()
    with pytest.raises(AssertionError):
        trainer(video, video_embed=emb)
    before1 = [p.detach().clone() for p in dec.unets[0].parameters()]
    before2 = [p.detach().clone() for p in dec.unets[1].parameters()]
    losses = []
    for _ in range(8):
        torch.manual_seed(5)
        losses.append(trainer(video, video_embed=emb, unet_number=2))
        trainer.update(unet_number=2)
    assert trainer.steps.tolist() == [0, 8]
    assert losses[-1] < 0.95 * losses[0], losses
    assert all(torch.equal(a, b) for a, b in zip(before1, dec.unets[0].parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(before2, dec.unets[1].parameters()))
    after2 = [p.detach().clone() for p in dec.unets[1].parameters()]
    trainer(video, video_embed=emb, unet_number=1)   # the first stage resizes the 32 px video down to its 16 px itself
    trainer.update(unet_number=1)
    assert trainer.steps.tolist() == [1, 8]
    assert all(torch.equal(a, b) for a, b in zip(after2, dec.unets[1].parameters()))
    assert any(not torch.equal(a, b) for a, b in zip(before1, dec.unets[0].parameters()))


@pytest.mark.parametrize("shape", [(3, 3, 3, 16, 32), (1, 3, 2, 8, 8), (2, 3, 5, 24, 40)])
def test_training_on_ragged_shapes_matches_oracle_autograd(shape):
    """Odd batch / frame counts, non-square frames and sides that are not a multiple of 16: the backward kernels' tiling
    fallbacks (position tails, images that do not fill a 128-row tile) against oracle autograd."""
    kw, _ = CASES["tiny"]
    dec, sd = build(kw, shape[2], shape[3])
    g = torch.Generator().manual_seed(sum(shape))
    video, noise = torch.rand(shape, generator=g), torch.randn(shape, generator=g)
    emb = torch.randn(shape[0], 512, generator=g)
    times = torch.randint(0, 100, (shape[0],), generator=g)
    ref_loss, ref_grads = oracle_loss_and_grads(sd, kw, video, times, noise, emb)
    unet = dec.unets[0]
    loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0],
                    noise=noise.cuda())
    assert abs(float(loss) - ref_loss) < 2e-2 * ref_loss, (float(loss), ref_loss)
    loss.backward()
    got = {k: p.grad for k, p in unet.named_parameters()}
    num = den = 0.0
    for k, rg in ref_grads.items():
        gg = got[k].float().cpu()
        num += float((gg - rg).pow(2).sum())
        den += float(rg.pow(2).sum())
        if rg.numel() >= 256 and float(rg.norm()) > 1e-6:
            assert float(F.cosine_similarity(gg.flatten(), rg.flatten(), dim=0)) >= 0.98, k
    assert (num / den) ** 0.5 < 8e-2, (num / den) ** 0.5


def test_training_with_scaled_skip_connections_matches_oracle_autograd(layouts):
    """scale_skip_connection=True (:926-928, :945): the 2^-0.5 factor on the skip half of every up-block concat is folded
    into the weight columns of that source for forward / data-gradient and restored on the weight gradient."""
    kw = dict(CASES["tiny"][0], scale_skip_connection=True)
    shape = (2, 3, 4, 16, 16)
    dec, sd = build(kw, shape[2], shape[3])
    g = torch.Generator().manual_seed(13)
    video, noise = torch.rand(shape, generator=g), torch.randn(shape, generator=g)
    emb, times = torch.randn(2, 512, generator=g), torch.tensor([11, 80])
    ref_loss, ref_grads = oracle_loss_and_grads(sd, kw, video, times, noise, emb)
    plain_loss, _ = oracle_loss_and_grads(sd, CASES["tiny"][0], video, times, noise, emb)
    assert abs(ref_loss - plain_loss) > 1e-3 * plain_loss   # the option changes the function
    unet = dec.unets[0]
    loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0],
                    noise=noise.cuda())
    assert abs(float(loss) - ref_loss) < 2e-2 * ref_loss, (float(loss), ref_loss)
    loss.backward()
    got = {k: p.grad for k, p in unet.named_parameters()}
    num = den = 0.0
    for k, rg in ref_grads.items():
        gg = got[k].float().cpu()
        num += float((gg - rg).pow(2).sum())
        den += float(rg.pow(2).sum())
        if rg.numel() >= 256 and float(rg.norm()) > 1e-6:
            assert float(F.cosine_similarity(gg.flatten(), rg.flatten(), dim=0)) >= 0.98, k
    assert (num / den) ** 0.5 < 8e-2, (num / den) ** 0.5


def test_eval_and_no_grad_forward_leave_the_gradient_buffer_untouched_and_chunks_run():
    """trainer.py:322-365: the backward runs only `if self.training`; `max_batch_size` splits the batch, each chunk's
    loss scaled by its share.  A validation pass must not leak gradients into the next update()."""
    kw, shape = CASES["tiny"]
    dec, _ = build(kw, shape[2], shape[3])
    trainer = VideoDecoderTrainer(dec, lr=3e-4)
    g = torch.Generator().manual_seed(9)
    video = torch.rand((4, *shape[1:]), generator=g).cuda()
    emb = torch.randn(4, 512, generator=g).cuda()
    trainer.eval()
    l_eval = trainer(video, video_embed=emb)
    assert l_eval > 0 and float(trainer.flat_g.abs().max()) == 0.0
    trainer.train()
    with torch.no_grad():
        trainer(video, video_embed=emb)
    assert float(trainer.flat_g.abs().max()) == 0.0
    before = trainer.flat_p.clone()
    trainer.update()                                   # all-zero gradient: Adam moves nothing but the decayed weights
    assert float((trainer.flat_p - before).abs().max()) < 5e-5   # lr * wd * |p| = 3e-4 * 1e-2 * 3.7
    dec.zero_grad()                                    # set_to_none=True detaches .grad from the flat buffer
    l_chunk = trainer(video, video_embed=emb, max_batch_size=3)   # chunks of 3 + 1: two plans, gradients accumulate
    assert l_chunk > 0 and float(trainer.flat_g.abs().max()) > 0
    w = dec.unets[0].to_time_cond[0].weight
    base = trainer.flat_g.data_ptr()
    assert base <= w.grad.data_ptr() < base + trainer.flat_g.numel() * 4
    trainer.update()
    assert float((trainer.flat_p - before).abs().max()) > 1e-4   # an Adam step moves every element by ~lr


def test_backward_through_a_plan_whose_buffers_were_overwritten_raises():
    """One plan per shape owns one set of activations: (l1 + l2).backward() over two forwards of the same shape would
    back-propagate l1 through l2's activations — it must fail loudly instead."""
    from dalle2_video_b200 import _lib as L
    kw, shape = CASES["tiny"]
    dec, _ = build(kw, shape[2], shape[3])
    g = torch.Generator().manual_seed(1)
    video, emb = torch.rand(shape, generator=g).cuda(), torch.randn(shape[0], 512, generator=g).cuda()
    l1 = dec(video, video_embed=emb)
    l2 = dec(video, video_embed=emb)
    with pytest.raises(L.DvError):
        l1.backward()
    l2.backward()                                      # the most recent forward is still valid
    # moving the parameters (a new optimiser flattening, .to(), .float()) rebuilds the plan instead of replaying graphs
    unet = dec.unets[0]
    plan_a = unet.train_plan(shape[0], shape[2], shape[3], shape[4], video.device)
    VideoDecoderTrainer(dec)
    plan_b = unet.train_plan(shape[0], shape[2], shape[3], shape[4], video.device)
    assert plan_a is not plan_b
