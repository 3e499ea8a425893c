"""Diagnostics for the tcgen05 projection kernel: error maps against fp64 for structured inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import numpy as np, torch
import velocity_asr as va

def run(M, K, N, kind):
    g = torch.Generator().manual_seed(1)
    if kind == "rand":
        x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5
    elif kind == "rowid":      # C[m, n] = m  (x row m = m at k=0, w[n,0] = 1)
        x = torch.zeros(M, K); x[:, 0] = torch.arange(M).float(); w = torch.zeros(N, K); w[:, 0] = 1
    elif kind == "colid":      # C[m, n] = n
        x = torch.zeros(M, K); x[:, 0] = 1; w = torch.zeros(N, K); w[:, 0] = torch.arange(N).float()
    elif kind == "kpos":       # C[m, n] = sum_k k * [k == n % K]
        x = torch.arange(K).float().repeat(M, 1); w = torch.zeros(N, K); w[torch.arange(N), torch.arange(N) % K] = 1
    y = va.linear(x.cuda(), w.cuda(), None, tensor_cores=True).cpu().double().numpy()
    ref = x.double().numpy() @ w.double().numpy().T
    err = np.abs(y - ref)
    scale = np.abs(ref).max() + 1e-30
    print(f"[{kind}] M={M} K={K} N={N}: max rel err {err.max()/scale:.3e}; frac wrong(>1e-4) {(err/scale > 1e-4).mean():.4f}")
    if err.max() / scale > 1e-4:
        bad = err / scale > 1e-4
        print("   wrong rows (first 16):", np.where(bad.any(1))[0][:16], " wrong cols (first 16):", np.where(bad.any(0))[0][:16])
        print("   y[0,:8]  ", y[0, :8], "\n   ref[0,:8]", ref[0, :8])
        print("   y[:8,0]  ", y[:8, 0], "\n   ref[:8,0]", ref[:8, 0])
        print("   y[1,:8]  ", y[1, :8], "\n   ref[1,:8]", ref[1, :8])

for kind in ("rowid", "colid", "kpos", "rand"):
    run(128, 32, 128, kind)
run(256, 64, 256, "rand")
run(1000, 192, 768, "rand")
run(130, 48, 200, "rand")
