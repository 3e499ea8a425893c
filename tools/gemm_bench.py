"""Micro-benchmark of the projection kernels at the model's shapes (config 2: M = 48,064 tokens).
REPS back-to-back launches between one pair of CUDA events (no host gap between launches); inputs rotate over
enough buffers to exceed L2."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va

SHAPES = [("in_proj", 48064, 192, 768, None, False), ("x_dt_proj", 48064, 384, 512, "softplus", False),
          ("out_proj", 48064, 384, 192, None, True), ("ffn1", 48064, 192, 384, "gelu", False),
          ("ffn2", 48064, 384, 192, None, True), ("ctc", 48064, 192, 1000, None, False),
          ("fusion3", 48064, 384, 576, None, False)]
REPS = 8

def run(name, M, K, N, act, resid, tc, iters=5):
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(M, K, device="cuda", generator=g) for _ in range(REPS)]
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    rs = [torch.randn(M, N, device="cuda", generator=g) for _ in range(REPS)] if resid and tc else [None] * REPS
    outs = [torch.empty(M, N, device="cuda") for _ in range(REPS)]
    ws = va.split_tf32(w) if tc else None
    def go():
        for i in range(REPS):
            va.linear(xs[i], w, b, activation=act, tensor_cores=tc, weight_split=ws, residual=rs[i], out=outs[i])
    go(); go()
    ms = []
    for _ in range(iters):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1) / REPS)
    ms.sort(); med = ms[len(ms) // 2]
    return {"name": name, "M": M, "K": K, "N": N, "resid": bool(resid and tc), "tc": tc, "ms": round(med, 4),
            "TFLOPs": round(2 * M * K * N / med / 1e9, 1)}

if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else None
    for s in SHAPES:
        if only and s[0] != only: continue
        for tc in ((True,) if only else (True, False)):
            print(json.dumps(run(*s, tc)), flush=True)
