"""BASELINE configs[3] in the reference's default scan mode: 16 x 600 s through the fused path, scan_mode="parallel"."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
torch.manual_seed(0)
m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="parallel"))
m.extend_positional_table(30008)
m = m.cuda().eval()
g = torch.Generator().manual_seed(1234)
audio = (torch.randn(16, 16000 * 600, generator=g) * 0.1).cuda()
m.transcribe(audio)
torch.cuda.synchronize()
ms = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.transcribe(audio); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print(json.dumps({"config": "16 x 600 s, scan_mode=parallel (reference default)", "ms": min(ms),
                  "rtfx": 16 * 600 / (min(ms) * 1e-3)}))
