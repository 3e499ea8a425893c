"""In-pipeline time of every projection launch of one config-2 step (CUDA events around each launch,
VASR_PROF=1), aggregated by shape: unlike the ncu launch list these are warm-cache, back-to-back times."""
import os, sys, ctypes
os.environ["VASR_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
from velocity_asr import _native
torch.manual_seed(0)
m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential")).cuda().eval()
audio = (torch.randn(64, 240000) * 0.1).cuda()
for _ in range(3): m.transcribe(audio)
eng = m._engines[0]
fn = eng.lib.vasr_debug_dump_profile; fn.restype = ctypes.c_int; fn.argtypes = [ctypes.c_void_p]
fn(eng.handle)                       # discard warm-up
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.transcribe(audio); e1.record(); torch.cuda.synchronize()
print("step (with event pairs) %.3f ms" % e0.elapsed_time(e1))
fn(eng.handle)
