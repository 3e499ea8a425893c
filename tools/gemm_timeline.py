"""Launch-level timeline of the CTA-pair projection kernel: per-CTA globaltimer stamps (entry, prologue done,
griddepcontrol.wait returned, first accumulator complete, first tile stored, last accumulator complete, epilogue
done, exit) of REPS back-to-back launches of one shape.  Shows where a launch's fixed cost goes.
    python tools/gemm_timeline.py M K N [resid]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
from velocity_asr import _native
lib = _native.lib()
fn = lib.vasr_debug_gemm_trace
fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
M, K, N = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (48064, 384, 192)
resid = len(sys.argv) > 4
REPS = 4
g = torch.Generator(device="cuda").manual_seed(1)
w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
ws = va.split_tf32(w)
xs = [torch.randn(M, K, device="cuda", generator=g) for _ in range(REPS)]
rs = [torch.randn(M, N, device="cuda", generator=g) for _ in range(REPS)]
outs = [torch.empty(M, N, device="cuda") for _ in range(REPS)]
def go(bufs=None):
    for i in range(REPS):
        if bufs is not None: fn(ctypes.c_void_p(bufs[i].data_ptr()))
        va.linear(xs[i], w, None, tensor_cores=True, weight_split=ws, residual=rs[i] if resid else None, out=outs[i])
go(); go(); torch.cuda.synchronize()
bufs = [torch.zeros(16 * 128 + 8 * 148, dtype=torch.int64, device="cuda") for _ in range(REPS)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); go(bufs); e1.record(); torch.cuda.synchronize(); fn(None)
print(f"M={M} K={K} N={N} resid={resid}: {e0.elapsed_time(e1) / REPS * 1e3:.1f} us per launch (events around {REPS} launches, trace on)")
names = ["entry", "wait ret", "1st acc", "1st tile out", "last acc", "epi done", "prologue", "exit"]
T = [b[16 * 128:].cpu().view(148, 8).double() for b in bufs]
t0 = min(float(t[:, 0][t[:, 0] > 0].min()) for t in T)
for i, t in enumerate(T):
    live = t[:, 0] > 0
    t = (t[live] - t0) / 1e3
    order = [0, 6, 1, 2, 3, 4, 5, 7]
    print(f"launch {i} ({int(live.sum())} CTAs), us since the first entry, min / median / max over CTAs:")
    for k in order:
        c = t[:, k]
        print(f"   {names[k]:13s} {float(c.min()):8.2f} {float(c.median()):8.2f} {float(c.max()):8.2f}")
