"""Pipeline trace of the tensor-core projection kernel: clock64 stamps of CTA 0 at
TMA issue (P), converter sees the tile (C0), converter done (C1), MMA warp released (M0), MMAs issued (M1)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
from velocity_asr import _native
lib = _native.lib()
lib._FuncPtr  # noqa
fn = lib.vasr_debug_gemm_trace
fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
M, K, N = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (48064, 192, 768)
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
ws = va.split_tf32(w)
for _ in range(3): va.linear(x, w, None, tensor_cores=True, weight_split=ws)
buf = torch.zeros(16 * 128 + 8 * 148, dtype=torch.int64, device="cuda")
fn(ctypes.c_void_p(buf.data_ptr()))
va.linear(x, w, None, tensor_cores=True, weight_split=ws)
torch.cuda.synchronize()
fn(None)
cta = buf[16 * 128:].cpu().view(148, 8)[:, [0, 7]]
t0 = int(cta[:, 0].min())
dur = [(int(c[1]) - int(c[0])) / 1e3 for c in cta]
start = [(int(c[0]) - t0) / 1e3 for c in cta]
end = [(int(c[1]) - t0) / 1e3 for c in cta]
print('per-CTA (us): start min/max %.1f/%.1f  duration min/med/max %.1f/%.1f/%.1f  last end %.1f' % (min(start), max(start), min(dur), sorted(dur)[74], max(dur), max(end)))
print('durations of CTAs 0..15:', [round(d, 1) for d in dur[:16]])
t = buf[:16 * 128].cpu().view(16, 128)
print('epilogue warp 6: per tile  wait-for-accumulator (E1-E0)  own-work (E0[i+1]-E1[i])  | MMA: tile period')
nk = (K + 31) // 32
for i in range(8):
    E0, E1, E0n = int(t[5, i]), int(t[6, i]), int(t[5, i + 1])
    print(f'  tile {i}: wait {E1 - E0:6d}  work {E0n - E1:6d}  (chunk 0: tcgen05.ld {int(t[7, i]) - E1:5d}, to smem {int(t[8, i]) - int(t[7, i]):5d}, read-back+store {int(t[9, i]) - int(t[8, i]):5d})   mma issue span {int(t[4, (i + 1) * nk - 1]) - int(t[3, i * nk]):6d}')
base = int(t[0, 0])
print("stage   P(issue)  C0(full)  C1(conv)  M0(go)  M1(issued) | tma=C0-P conv=C1-C0 wake=M0-C1 issue=M1-M0 period=P[i]-P[i-1]")
import statistics
per = [int(t[0, i]) - int(t[0, i - 1]) for i in range(8, 100)]
print('median period', statistics.median(per), 'mean', sum(per) / len(per), 'kernel span (CTA 0, first 128 stages)', int(t[4, 127]) - base)
for i in range(int(os.environ.get('ROWS', '40'))):
    P, C0, C1, M0, M1 = [int(t[r, i]) - base for r in range(5)]
    prev = int(t[0, i - 1]) - base if i else 0
    print(f"{i:3d} {P:9d} {C0:9d} {C1:9d} {M0:8d} {M1:9d} | {C0-P:6d} {C1-C0:6d} {M0-C1:6d} {M1-M0:6d} {P-prev:6d}")

P = [int(t[0, i]) for i in range(128)]
print("producer: clocks between consecutive stage releases (= completion cadence of the k-blocks' MMAs), k-blocks 24..95, one tile per line")
for r in range(24, 96, nk):
    print("   ", " ".join(f"{P[i] - P[i - 1]:5d}" for i in range(r, r + nk)), "  tile:", P[r + nk - 1] - P[r - 1])
if int(t[10, 1]) == 0:
    sys.exit(0)
print("MMA warp at tile boundaries: M1(last kb of tile i-1) -> loop top (T0) -> accumulator free (T1) -> M0(first kb)")
for i in range(1, 9):
    last = i * nk - 1
    print(f"  tile {i}: M1[{last}]..T0 {int(t[10, i]) - int(t[4, last]):6d}   T0..T1 {int(t[11, i]) - int(t[10, i]):6d}   T1..M0 {int(t[3, last + 1]) - int(t[11, i]):6d}")
