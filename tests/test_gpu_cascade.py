"""-m gpu: cascade stage 2 through the C ABI — the per-frame nearest resize, the noise augmentation
of the conditioning video, a lowres-conditioned Unet3D forward and VideoDecoder.sample over a two-unet cascade, against
fixtures generated from the reference's own code (oracle/make_golden.py cascade) with the noise draws shared by
construction.  Bars as test_gpu_unet / test_gpu_sampler: fp32 1e-4 relative per forward, 2e-3 absolute on the final sample;
bf16 3e-2 per forward, PSNR >= 30 dB on the final sample."""
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from dalle2_video_b200 import Unet3D, VideoDecoder
from dalle2_video_b200 import _lib as L
from dalle2_video_b200.decoder import LowresVideoConditioner, resize_video_to
from gpu_util import psnr, rel_err
from oracle import primitives as P
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hin,hout", [(16, 32), (64, 128), (32, 12), (10, 24), (7, 7)])
@pytest.mark.parametrize("clamp", [None, (0.0, 1.0)])
def test_resize_nearest_matches_interpolate(hin, hout, clamp):
    g = torch.Generator().manual_seed(hin * 131 + hout)
    v = (torch.rand(2, 3, 5, hin, hin, generator=g) * 1.6 - 0.3).cuda()
    out = resize_video_to(v, hout, clamp_range=clamp)
    if hin == hout:
        assert out is v
        return
    ref = F.interpolate(v.cpu().transpose(1, 2).reshape(10, 3, hin, hin), hout, mode="nearest")
    ref = ref.reshape(2, 5, 3, hout, hout).transpose(1, 2)
    if clamp is not None:
        ref = ref.clamp(*clamp)
    assert torch.equal(out.cpu(), ref)   # a gather: bit-exact


def test_noise_video_matches_q_sample():
    cond = LowresVideoConditioner(use_blur=False, use_noise=True, input_video_range=(0.0, 1.0), normalize_video=True)
    g = torch.Generator().manual_seed(3)
    v, nz = torch.rand(3, 3, 4, 8, 8, generator=g), torch.randn(3, 3, 4, 8, 8, generator=g)
    levels = torch.tensor([0, 200, 999])
    out, lv = cond.noise_video(v.cuda(), levels.cuda(), noise=nz.cuda())
    s = P.NoiseScheduler(beta_schedule="linear", timesteps=1000, loss_type="l2")
    ref = (s.q_sample(v * 2 - 1, t=levels, noise=nz) + 1) * 0.5
    assert rel_err(out, ref) < 1e-6 and torch.equal(lv.cpu(), levels)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_to_out_with_lowres_planes(dtype):
    B, P_, C, Ce = 2, 300, 16, 3
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B * P_, C, generator=g).to(dtype).cuda()
    extra = torch.randn(B, Ce, P_, generator=g).cuda()
    w, b = torch.randn(3, C + Ce, generator=g).cuda(), torch.randn(3, generator=g).cuda()
    out = torch.empty(B, 3, P_, device="cuda")
    L.call("dv_to_out", L.ToOutParams(x=x.data_ptr(), w=w.data_ptr(), bias=b.data_ptr(), out=out.data_ptr(), B=B, P=P_, C=C,
                                      Cout=3, dtype=L.dtype_code(dtype), extra=extra.data_ptr(), Cextra=Ce),
           torch.cuda.current_stream().cuda_stream)
    cat = torch.cat((x.float().reshape(B, P_, C), extra.transpose(1, 2)), dim=-1)
    ref = (cat @ w.t() + b).transpose(1, 2)
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_lowres_conditioned_forward_matches_reference_fixture(dtype, tol, layouts):
    g = golden("unet3d_lowres.pt")
    u = Unet3D(**g["unet_kwargs"])
    u.load_state_dict(synth_state_dict(layouts["cascade_u2_noise"], seed=g["weight_seed"]), strict=True)
    u = u.cuda()
    u.compute_dtype = dtype
    out = u(g["x"].cuda(), g["times"].cuda(), video_embed=g["video_embed"].cuda(),
            lowres_cond_video=g["lowres_cond_video"].cuda(), lowres_noise_level=g["lowres_noise_level"].cuda())
    assert out.shape == g["eps"].shape
    err = rel_err(out, g["eps"])
    assert err < tol, err
    # the conditioning video matters (guards against a plan that silently ignores it) and is required
    out2 = u(g["x"].cuda(), g["times"].cuda(), video_embed=g["video_embed"].cuda(),
             lowres_cond_video=torch.zeros_like(g["x"]).cuda(), lowres_noise_level=g["lowres_noise_level"].cuda())
    assert rel_err(out2, g["eps"]) > 10 * tol
    with pytest.raises(AssertionError):
        u(g["x"].cuda(), g["times"].cuda(), video_embed=g["video_embed"].cuda())


def shared_noise(seed):
    """noise_fn drawing every tensor from ONE CPU generator in call order: stage-1 loop, [augmentation], stage-2 loop —
    the reference's order under a single torch.manual_seed (:1689, :1657, :2145)."""
    gen = torch.Generator().manual_seed(seed)
    return lambda shape, i: torch.randn(shape, generator=gen)


def cascade_decoder(g, layouts):
    k1, k2 = g["unet_kwargs"]
    u1, u2 = Unet3D(**k1), Unet3D(**k2)
    d = VideoDecoder((u1, u2), frame_sizes=g["frame_sizes"], frame_numbers=(g["frames"],) * 2, timesteps=g["timesteps"],
                     learned_variance=False, use_noise_for_lowres_cond=g["use_noise"])
    assert d.unets[1].lowres_cond and d.unets[1].lowres_noise_cond == g["use_noise"]
    d.unets[0].load_state_dict(synth_state_dict(layouts["tiny"], seed=g["weight_seeds"][0]), strict=True)
    key = "cascade_u2_noise" if g["use_noise"] else "cascade_u2"
    d.unets[1].load_state_dict(synth_state_dict(layouts[key], seed=g["weight_seeds"][1]), strict=True)
    return d.cuda()


@pytest.mark.parametrize("kind", ["plain", "noise"])
@pytest.mark.parametrize("graph", [False, True])
def test_cascade_sample_fp32_matches_reference_fixture(kind, graph, layouts):
    g = golden(f"sample_cascade_{kind}.pt")
    d = cascade_decoder(g, layouts)
    d.use_cuda_graph = graph
    emb = g["video_embed"].cuda()
    s1 = d.sample(video_embed=emb, stop_at_unet_number=1, noise_fn=shared_noise(g["sample_seed"]))
    assert float((s1.cpu() - g["stage1"]).abs().max()) < 2e-3
    vid = d.sample(video_embed=emb, noise_fn=shared_noise(g["sample_seed"]))
    assert vid.shape == g["sample"].shape
    err = float((vid.cpu() - g["sample"]).abs().max())
    assert err < 2e-3, err
    # start_at_unet_number=2 from the stage-1 video resumes the cascade (:2086-2095)
    gen = torch.Generator().manual_seed(g["sample_seed"])
    for _ in range(g["timesteps"] + 1):
        torch.randn(g["stage1"].shape, generator=gen)          # skip the stage-1 draws
    vid2 = d.sample(video=g["stage1"].cuda(), video_embed=emb, start_at_unet_number=2,
                    noise_fn=lambda shape, i: torch.randn(shape, generator=gen))
    assert float((vid2.cpu() - g["sample"]).abs().max()) < 2e-3


def test_cascade_sample_bf16_psnr_bound(layouts):
    g = golden("sample_cascade_noise.pt")
    d = cascade_decoder(g, layouts)
    for u in d.unets:
        u.compute_dtype = torch.bfloat16
    vid = d.sample(video_embed=g["video_embed"].cuda(), noise_fn=shared_noise(g["sample_seed"]))
    p = psnr(vid, g["sample"])
    print(f"bf16 cascade PSNR vs fp32 reference: {p:.1f} dB")
    assert p >= 30.0, p
