"""A/B of two builds on one box: device-resident step time (vasr_transcribe, config 2) for the library VASR_LIB names.
    VASR_LIB=.../libvasr_base.so python tools/step_ab.py ; python tools/step_ab.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
from velocity_asr import _native

mode = sys.argv[1] if len(sys.argv) > 1 else "sequential"
B, S = 64, 240000
torch.manual_seed(0)
m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode=mode)).cuda().eval()
eng = m._engine(torch.device("cuda", 0))
g = torch.Generator().manual_seed(1234)
pcm = [(torch.randn(B, S, generator=g) * 0.1).cuda() for _ in range(4)]
L = (1 + S // 160 + 1) // 2
tok = torch.empty(B, L, dtype=torch.int32, device="cuda"); ln = torch.empty(B, dtype=torch.int32, device="cuda")
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def step(i):
    _native.check(eng.lib.vasr_transcribe(eng.handle, _native.ptr(pcm[i % 4]), B, S, _native.ptr(tok), _native.ptr(ln), sp))
for i in range(5): step(i)
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): step(i)
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 20)
print(os.environ.get("VASR_LIB", "libvasr.so"), mode, "ms/step", [round(r, 3) for r in res])
