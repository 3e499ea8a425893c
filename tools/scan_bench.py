"""Micro-benchmark of the selective-scan operator at BASELINE config 2 shape (64 x 751 x 384, N=64).
Prints ms per launch and achieved algorithmic GB/s (6,656 B per token-layer with the gate)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va

def run(B=64, L=751, Di=384, N=64, mode="sequential", structured=True, gate=True, iters=20, dt_scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, L, Di, device="cuda", generator=g)
    dt = torch.nn.functional.softplus(torch.randn(B, L, Di, device="cuda", generator=g)) * dt_scale
    z = torch.randn(B, L, Di, device="cuda", generator=g) if gate else None
    Bm = torch.randn(B, L, N, device="cuda", generator=g)
    Cm = torch.randn(B, L, N, device="cuda", generator=g)
    A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32)
    if not structured:
        A = A * 1.0371
    D = torch.randn(Di, device="cuda", generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        va.selective_scan(x, dt, A, Bm, Cm, D, z=z, scan_mode=mode, validate=False)
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); va.selective_scan(x, dt, A, Bm, Cm, D, z=z, scan_mode=mode, validate=False); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    ms.sort(); med = ms[len(ms) // 2]
    byt = B * L * 4 * ((4 if gate else 3) * Di + 2 * N)
    return {"B": B, "L": L, "N": N, "mode": mode, "structured": structured, "dt_scale": dt_scale, "ms": round(med, 4),
            "GBps": round(byt / med / 1e6, 1), "frac_6545": round(byt / med / 1e6 / 6545.6, 4)}

if __name__ == "__main__":
    quick = "--quick" in sys.argv
    if "--L" in sys.argv:        # length sweep in parallel mode: --L 128 256 751
        i = sys.argv.index("--L")
        for l in sys.argv[i + 1:]:
            print(json.dumps(run(L=int(l), mode="parallel", iters=10)), flush=True)
        sys.exit(0)
    if "--B" in sys.argv:        # batch sweep: --B 49 64 74 ...
        i = sys.argv.index("--B")
        mode = "parallel" if "--quirk" in sys.argv else "sequential"
        for b in sys.argv[i + 1:]:
            print(json.dumps(run(B=int(b), mode=mode, iters=10)), flush=True)
        sys.exit(0)
    if "--quirk" in sys.argv:
        # random-init scale of dt (the decay reaches an exact zero after ~150 steps) and a small-dt model (never)
        cases = [dict(mode="parallel", N=32, L=93), dict(mode="parallel"), dict(mode="parallel", dt_scale=0.02)]
    else:
      cases = [dict()] if quick else [dict(), dict(structured=False), dict(mode="parallel"), dict(N=32, L=93),
                                    dict(B=16, L=30001, iters=5), dict(B=512, iters=5), dict(B=1, L=501)]
    for c in cases:
        print(json.dumps(run(**c)), flush=True)
