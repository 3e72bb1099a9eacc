"""Resnet-block outputs of the float32 training plan against the float32 sampler engine, block by block (debug aid)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_train import CASES, build
from dalle2_video_b200 import unet3d as U, train as TR, _lib as L
from dalle2_video_b200.train import p_losses
from oracle.primitives import NoiseScheduler as OracleScheduler
kw, shape = CASES["headline_small"]
kw = dict(kw, dim=int(sys.argv[1]), dim_mults=tuple(int(v) for v in sys.argv[2].split(",")))
shape = tuple(int(v) for v in sys.argv[3].split(","))
dec, sd = build(kw, shape[2], shape[3])
g = torch.Generator().manual_seed(3)
video = torch.rand(shape, generator=g); noise = torch.randn(shape, generator=g); emb = torch.randn(shape[0], 512, generator=g)
times = torch.tensor([7, 63])
sched = OracleScheduler(beta_schedule="cosine", timesteps=100, loss_type="l2")
x_t = sched.q_sample(2 * video - 1, times, noise)
unet = dec.unets[0]
unet.compute_dtype = torch.float32

eng_marks, plan_marks = [], []
orig_e = U._Engine._resblock
def rb_e(self, m, srcs, h, w, skip_scale=None):
    out = orig_e(self, m, srcs, h, w, skip_scale)
    eng_marks.append((len(self.ops), out, srcs[0].C))
    return out
U._Engine._resblock = rb_e
orig_p = TR._TrainPlan._resblock
def rb_p(self, m, srcs, h, w, skip_scale=1.0):
    out = orig_p(self, m, srcs, h, w, skip_scale)
    plan_marks.append(out)
    return out
TR._TrainPlan._resblock = rb_p

eng = unet.engine(shape[0], shape[2], shape[3], shape[4], torch.float32, torch.device("cuda:0"), 0)
keep = torch.ones(shape[0], dtype=torch.uint8, device="cuda")
eng.set_lowres(None, None)
eng.set_conditioning(emb.cuda(), None, keep, keep)
eng.x_in.copy_(x_t.cuda()); eng.time_in.copy_(times.cuda()); eng.set_time_source(False)
stream = torch.cuda.current_stream().cuda_stream
eng.stats.zero_()
snaps, mi = [], 0
for i, (name, p) in enumerate(eng.ops):
    L.call(name, p, stream)
    while mi < len(eng_marks) and eng_marks[mi][0] == i + 1:
        act = eng_marks[mi][1]
        torch.cuda.synchronize()
        snaps.append(act.buf.view(torch.float32)[: act.buf.numel() // 4].clone())
        mi += 1
loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0], noise=noise.cuda())
torch.cuda.synchronize()
print(len(snaps), len(plan_marks))
for i, (s, v) in enumerate(zip(snaps, plan_marks)):
    a = v.t.float().flatten(); b = s[: a.numel()]
    print(i, tuple(v.t.shape), "rel", float((a - b).norm() / b.norm()))

# ---- op-by-op: outputs of every op with an `out` field, engine (snapshotted as it runs) vs plan
def out_size(name, p):
    if name == "dv_conv_igemm":
        ho, wo = (p.H // 2, p.W // 2) if p.mode == L.CONV_UNSHUFFLE else (p.H, p.W)
        ld = p.out_ld if p.out_ld else (p.Cout // 4 if p.epilogue != L.EPI_STORE else p.Cout)
        rows = p.n_img * ho * wo * (4 if p.epilogue != L.EPI_STORE else 1)
        return rows * ld
    if name == "dv_gn_apply": return p.B * p.P * p.C
    if name == "dv_layernorm": return p.rows * p.C
    if name == "dv_xattn_core": return p.B * p.P * p.heads * p.dh
    if name == "dv_attn_mqa": return p.S * p.n * p.heads * p.dh
    return None
def from_ptr(ptr, n):
    import ctypes
    t = torch.empty(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    global _rt
    try: _rt
    except NameError:
        import glob
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + ["libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"]
        for c in cands:
            try: _rt = ctypes.CDLL(c); break
            except OSError: pass
    rc = _rt.cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(n * 4), 3)
    assert rc == 0, rc
    return t
eng.stats.zero_()
E_out = []
for name, p in eng.ops:
    L.call(name, p, stream)
    n = out_size(name, p)
    if n: E_out.append((name, n, from_ptr(p.out, n)))
plan = unet._train_plan_last
P_out = []
for name, p in plan.fwd:
    n = out_size(name, p)
    if n: P_out.append((name, n, from_ptr(p.out, n)))
E_out = E_out[3:]; P_out = P_out[1:]
for i in range(min(len(E_out), len(P_out))):
    (en, es, et), (pn, ps, pt) = E_out[i], P_out[i]
    if en != pn or es != ps:
        print(i, "misaligned", en, es, pn, ps); break
    r = float((et - pt).norm() / et.norm().clamp_min(1e-30))
    if r > 1e-4 or 40 <= i <= 75: print(i, en, es, "rel", r)
ei = [i for i, (n, p) in enumerate(eng.ops) if out_size(n, p)][3 + 59]
pi = [i for i, (n, p) in enumerate(plan.fwd) if out_size(n, p)][1 + 59]
for tag, (n, p) in (("engine", eng.ops[ei]), ("plan", plan.fwd[pi])):
    print(tag, n, {f[0]: getattr(p, f[0]) for f in p._fields_ if f[0] not in ("src0", "src1", "weight", "bias", "out", "stats", "add")})
en, ep = eng.ops[ei][1], plan.fwd[pi][1]
ktot = 9 * en.C0
we = from_ptr(en.weight, en.Cout * ktot); wp = from_ptr(ep.weight, ep.Cout * ktot)
print("weights rel", float((we - wp).norm() / we.norm()))
be = from_ptr(en.bias, en.Cout); bp = from_ptr(ep.bias, ep.Cout)
print("bias rel", float((be - bp).norm() / be.norm().clamp_min(1e-30)))
xe = from_ptr(en.src0, 128 * en.C0); xp = from_ptr(ep.src0, 128 * ep.C0)
print("input (final buffer contents) rel", float((xe - xp).norm() / xe.norm()))
tw, fn = [(t, f) for t, f in plan.refresh if t.data_ptr() == ep.weight][0]
with torch.no_grad():
    good = fn().to(tw.dtype)
print("weight tensor vs its producer now:", float((tw - good).norm() / good.norm()), tuple(tw.shape))
plan.refresh_weights(); torch.cuda.synchronize()
print("after refresh:", float((tw - good).norm() / good.norm()))
plan.stats.zero_()
for i, (name, p) in enumerate(plan.fwd):
    L.call(name, p, stream); torch.cuda.synchronize()
    e = float((tw - good).norm() / good.norm())
    if e > 0:
        print("corrupted after fwd op", i, name, {f[0]: getattr(p, f[0]) for f in p._fields_}, "err", e)
        break
