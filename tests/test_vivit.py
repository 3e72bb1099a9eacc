"""Scope row N4: the ViViT video-encoder forward.  CPU: the oracle restatement against fixtures
generated from the reference's own module, state_dict layout and same-seed initialisation of the drop-in.  GPU (-m gpu):
the native forward (dv_patchify / dv_linear_small / dv_layernorm / dv_mha_small) against the same fixtures, float32,
1e-4 relative (the north_star bar for fp32)."""
import importlib.util
import os

import pytest
import torch

from conftest import golden
from dalle2_video_b200.vivit import ViViT
from oracle import reference
from oracle.make_golden import vivit_fixture_inputs
from oracle.vivit import vivit_forward

CASES = ["small", "dim512"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    g = golden(f"vivit_{name}.pt")
    sd, x = vivit_fixture_inputs(g)
    kw = g["kwargs"]
    with torch.no_grad():
        out = vivit_forward(sd, x, patch_size=kw["patch_size"], depth=kw["depth"], heads=kw["heads"])
    rel = float((out - g["out"]).norm() / g["out"].norm())
    assert rel < 1e-5, rel


@pytest.mark.parametrize("name", CASES)
def test_state_dict_layout_matches_reference(name):
    g = golden(f"vivit_{name}.pt")
    enc = ViViT(**g["kwargs"])
    assert {k: list(v.shape) for k, v in enc.state_dict().items()} == g["layout"]
    assert list(enc.state_dict()) == list(g["layout"])


@pytest.mark.skipif(not reference.available(), reason="/root/reference not mounted")
def test_same_seed_gives_reference_init():
    path = os.path.join(reference.REFERENCE_ROOT, "dalle2_video", "video_encoder.py")
    spec = importlib.util.spec_from_file_location("_reference_video_encoder_t", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    kw = dict(frame_size=32, patch_size=8, num_frames=25, dim=64, depth=2, heads=3, in_channels=3)
    torch.manual_seed(3)
    a = mod.ViViT(**kw).state_dict()
    torch.manual_seed(3)
    b = ViViT(**kw).state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_cpu_tensors_raise_loudly():
    enc = ViViT(frame_size=32, patch_size=8, num_frames=25, dim=64, depth=1)
    with pytest.raises(Exception, match="no CPU path"):
        enc(torch.zeros(1, 25, 3, 32, 32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_native_forward_matches_reference_fixture(name):
    g = golden(f"vivit_{name}.pt")
    sd, x = vivit_fixture_inputs(g)
    enc = ViViT(**g["kwargs"])
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda()
    out = enc(x.cuda())
    assert out.shape == g["out"].shape
    rel = float((out.cpu() - g["out"]).norm() / g["out"].norm())
    assert rel < 1e-4, rel
    # batch rows are independent (the encoder shards by video like the sampler)
    if x.shape[0] > 1:
        one = enc(x[:1].cuda())
        assert float((one.cpu() - out[:1].cpu()).abs().max()) < 1e-5


@pytest.mark.gpu
def test_patchify_and_small_mha_kernels():
    from dalle2_video_b200 import _lib as L
    g = torch.Generator().manual_seed(0)
    x = torch.randn(6, 3, 16, 24, generator=g).cuda()
    p = 4
    out = torch.empty(6 * 4 * 6, p * p * 3, device="cuda")
    L.call("dv_patchify", L.PatchifyParams(x=x.data_ptr(), out=out.data_ptr(), frames=6, C=3, H=16, W=24, patch=p),
           torch.cuda.current_stream().cuda_stream)
    ref = x.reshape(6, 3, 4, p, 6, p).permute(0, 2, 4, 3, 5, 1).reshape(6 * 24, p * p * 3)
    assert torch.equal(out, ref)
    S, n, H, D = 3, 150, 3, 64
    qkv = torch.randn(S, n, 3 * H * D, generator=g).cuda()
    o = torch.empty(S, n, H * D, device="cuda")
    L.call("dv_mha_small", L.MhaSmallParams(qkv=qkv.data_ptr(), out=o.data_ptr(), S=S, n=n, heads=H, dh=D, scale=D ** -0.5),
           torch.cuda.current_stream().cuda_stream)
    q, k, v = (t.reshape(S, n, H, D).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    ref = ((q @ k.transpose(-1, -2)) * D ** -0.5).softmax(-1) @ v
    ref = ref.permute(0, 2, 1, 3).reshape(S, n, H * D)
    assert float((o - ref).norm() / ref.norm()) < 1e-5
