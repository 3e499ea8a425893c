"""Micro-benchmark of the projection kernels at the model's shapes (config 2: M = 48,064 tokens)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va

SHAPES = [("in_proj", 48064, 192, 768, None), ("x_dt_proj", 48064, 384, 512, "softplus"),
          ("out_proj", 48064, 384, 192, None), ("ffn1", 48064, 192, 384, "gelu"), ("ffn2", 48064, 384, 192, None),
          ("ctc", 48064, 192, 1000, None), ("fusion3", 48064, 384, 576, None)]

def run(name, M, K, N, act, tc, iters=10):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(M, K, device="cuda", generator=g); w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    ws = va.split_tf32(w) if tc else None
    for _ in range(2): va.linear(x, w, b, activation=act, tensor_cores=tc, weight_split=ws)
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); va.linear(x, w, b, activation=act, tensor_cores=tc, weight_split=ws); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    ms.sort(); med = ms[len(ms) // 2]
    return {"name": name, "M": M, "K": K, "N": N, "tc": tc, "ms": round(med, 4), "TFLOPs": round(2 * M * K * N / med / 1e9, 1)}

if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else None
    for s in SHAPES:
        if only and s[0] != only: continue
        for tc in ((True,) if only else (True, False)):
            print(json.dumps(run(*s, tc)), flush=True)
