"""Time the device prefix beam search at the headline shape (64 x 751 frames x vocab 1000, width 10).

    python tools/beam_bench.py [--width 10] [--batch 64]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "velocity-asr_b200"))
import velocity_asr as va  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=10)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=751)
ap.add_argument("--vocab", type=int, default=1000)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
lg = torch.randn(a.batch, a.frames, a.vocab, device="cuda", generator=g) * 3.0
for _ in range(2):
    va.ctc_beam_search(lg, beam_width=a.width)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    res = va.ctc_beam_search(lg, beam_width=a.width)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
import json
print(json.dumps({"what": "ctc_beam_search on the device, host result lists included", "batch": a.batch,
                  "frames": a.frames, "vocab": a.vocab, "beam_width": a.width, "ms_per_batch": round(dt * 1e3, 3),
                  "best_beam_len": len(res[0][0].tokens)}))
