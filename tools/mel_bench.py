"""Time the one-pass log-mel front end at config 2 (64 x 15 s): mel_fft + stats combine + finish."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
pcm = [torch.randn(64, 240000, device="cuda") * 0.1 for _ in range(4)]
for p in pcm: va.compute_mel_spectrogram(p)
ms = []
for i in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); va.compute_mel_spectrogram(pcm[i % 4]); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
ms.sort()
print("compute_mel_spectrogram 64 x 15 s: %.3f ms (median of 12)" % ms[len(ms) // 2])
