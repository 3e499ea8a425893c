"""Where the end-to-end time of VELOCITYASR.transcribe(host tensor) goes."""
import os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
from velocity_asr import _native
torch.manual_seed(0)
m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential")).cuda().eval()
B, S = 64, 240000
host = (torch.randn(B, S) * 0.1).pin_memory()
for _ in range(3): m.transcribe(host)
def t(f, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("transcribe(host) ms", t(lambda: m.transcribe(host)))
dev = torch.device("cuda", 0)
print("_engine() ms", t(lambda: m._engine(dev)))
eng = m._engine(dev)
L = 751
tok = torch.empty(B, L, dtype=torch.int32, pin_memory=True); ln = torch.empty(B, dtype=torch.int32, pin_memory=True)
print("vasr_transcribe_host ms", t(lambda: eng.lib.vasr_transcribe_host(eng.handle, _native.ptr(host), B, S, _native.ptr(tok), _native.ptr(ln))))
d = host.cuda(); tk = torch.empty(B, L, dtype=torch.int32, device="cuda"); l2 = torch.empty(B, dtype=torch.int32, device="cuda")
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("vasr_transcribe(dev) ms", t(lambda: eng.lib.vasr_transcribe(eng.handle, _native.ptr(d), B, S, _native.ptr(tk), _native.ptr(l2), sp)))
print("H2D copy ms", t(lambda: d.copy_(host, non_blocking=True)))
print("list building ms", t(lambda: [tok[b, : int(ln[b])].tolist() for b in range(B)]))
print("pinned alloc ms", t(lambda: torch.empty(B, L, dtype=torch.int32, pin_memory=True)))
