"""BASELINE configs[3]: 16 x 600 s utterances (60,001 frames, 30,001 tokens) through the fused path.
Needs pe_time regenerated to >= 30,001 rows (the reference stops at 5,000, model.py:87,125)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
B, SEC = int(os.environ.get("LF_B", 16)), int(os.environ.get("LF_SEC", 600))
torch.manual_seed(0)
m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential"))
m.extend_positional_table(SEC * 50 + 8)
m = m.cuda().eval()
g = torch.Generator().manual_seed(1234)
audio = (torch.randn(B, 16000 * SEC, generator=g) * 0.1).cuda()
out = m.transcribe(audio)             # warm-up (workspace allocation)
torch.cuda.synchronize()
ms = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = m.transcribe(audio); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
# an utterance decodes the same alone as inside the batch (sharding rule) -- also at this length
one = m.transcribe(audio[1:2].contiguous())
print(json.dumps({"config": f"{B} x {SEC} s", "tokens_per_utt": (1 + 16000 * SEC // 160 + 1) // 2, "ms": min(ms),
                  "rtfx": B * SEC / (min(ms) * 1e-3), "decoded_tokens": [len(t) for t in out][:4],
                  "alone_equals_batched": one[0] == out[1],
                  "workspace_GB": m._engines[0].lib.vasr_workspace_bytes(m._engines[0].handle) / 1e9}))
