"""Per-kernel-family time of one config-2 step from torch.profiler-free CUDA events is not available for the
non-projection kernels, so this prints the device time of N back-to-back steps (quick A/B of small kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch, ctypes
import velocity_asr as va
from velocity_asr import _native
torch.manual_seed(0)
m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential")).cuda().eval()
audios = [(torch.randn(64, 240000) * 0.1).cuda() for _ in range(4)]
for a in audios: m.transcribe(a)
eng = m._engines[0]
tok = torch.empty(64, 751, dtype=torch.int32, device="cuda"); ln = torch.empty(64, dtype=torch.int32, device="cuda")
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(20):
        _native.check(eng.lib.vasr_transcribe(eng.handle, _native.ptr(audios[i % 4]), 64, 240000, _native.ptr(tok), _native.ptr(ln), sp))
    e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 20)
print("ms per step %.4f" % best)
