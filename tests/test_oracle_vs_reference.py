"""CPU, only where /root/reference is mounted: the functional restatement (oracle/unet3d.py) against the
reference's own modules imported from where they lie, on configurations the fixtures do not cover."""
import pytest
import torch

from oracle import reference
from oracle.unet3d import UnetConfig, forward_with_cond_scale
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.skipif(not reference.available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("kw,shape", [
    (dict(dim=32, video_embed_dim=512, dim_mults=(1, 2, 2), cond_on_video_embeds=True, num_resnet_blocks=1), (1, 3, 2, 16, 16)),
    (dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True, self_attn=(False, True)), (1, 3, 2, 16, 16)),
    (dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=False), (2, 3, 2, 16, 16)),
    (dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True, scale_skip_connection=True,
          self_cond=True), (1, 3, 2, 16, 16)),
])
def test_restatement_equals_reference_modules(kw, shape):
    ref = reference.load()
    torch.manual_seed(0)
    unet = ref.Unet3D(**kw).eval()
    sd = synth_state_dict({k: list(v.shape) for k, v in unet.state_dict().items()}, seed=3)
    unet.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g)
    emb = torch.randn(shape[0], 512, generator=g)
    t = torch.randint(0, 1000, (shape[0],), generator=g)
    sc = torch.randn(shape, generator=g) if kw.get("self_cond") else None
    with torch.no_grad():
        a = unet.forward_with_cond_scale(x, t, video_embed=emb, cond_scale=1.5 if kw.get("cond_on_video_embeds") else 1.0,
                                         self_cond=sc)
        b = forward_with_cond_scale(sd, UnetConfig(**kw), x, t, video_embed=emb,
                                    cond_scale=1.5 if kw.get("cond_on_video_embeds") else 1.0, self_cond=sc)
    assert float((a - b).norm() / a.norm()) < 1e-5


def test_zero_init_to_out_gives_zero_and_sampler_mean():
    # anchor (ii) of SURVEY.md §8c: a freshly constructed reference Unet3D outputs exactly 0
    ref = reference.load()
    torch.manual_seed(0)
    unet = ref.Unet3D(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True).eval()
    with torch.no_grad():
        y = unet(torch.randn(1, 3, 2, 16, 16), torch.tensor([5]), video_embed=torch.randn(1, 512))
    assert float(y.abs().max()) == 0.0
