"""Generate tests/golden/beam.npz by RUNNING the unmodified reference's ctc_beam_search (decode.py:128-217).

    python tests/golden/make_golden_beam.py          (build container: /root/reference importable)

Cases: plain random logits; logits on a coarse grid with few classes (many exactly tied scores, so the
reference's insertion-order tie-breaks decide the ranking); a vocabulary smaller than the beam; a blank
that is not token 0; peaked frames with repeats across blanks; beam_width 1.
Stored per case i: lg_i (B, L, V) fp32, par_i = (beam_width, blank), tok_i (B, W, L) int32 (-1 padded),
len_i (B, W) (-1 = fewer beams than W), sc_i (B, W) fp64, lp_i = the reference's fp32 log_softmax table.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refload import load_reference  # noqa: E402

R = load_reference()
assert R is not None, "reference not importable"
rs = np.random.RandomState(11)


def cases():
    yield rs.standard_normal((3, 40, 30)).astype(np.float32) * 2.0, 5, 0
    g = np.round(rs.standard_normal((2, 25, 6)) * 2.0) * 0.5
    yield g.astype(np.float32), 4, 0
    yield rs.standard_normal((2, 12, 3)).astype(np.float32), 10, 0
    yield rs.standard_normal((2, 30, 12)).astype(np.float32) * 3.0, 6, 2
    pk = np.full((1, 16, 9), -4.0, np.float32)
    for i, t in enumerate([0, 5, 5, 0, 5, 7, 7, 7, 0, 0, 3, 3, 5, 0, 1, 1]):
        pk[0, i, t] = 4.0
    yield pk + 0.3 * rs.standard_normal(pk.shape).astype(np.float32), 8, 0
    yield rs.standard_normal((2, 20, 16)).astype(np.float32), 1, 0
    yield np.zeros((1, 6, 5), np.float32), 3, 0          # every score tied: pure insertion order


out = {}
n = 0
for lg, W, blank in cases():
    res = R.decode.ctc_beam_search(torch.from_numpy(lg), beam_width=W, blank_token=blank)
    B, L, _ = lg.shape
    tok = np.full((B, W, L), -1, np.int32)
    ln = np.full((B, W), -1, np.int32)
    sc = np.full((B, W), -np.inf, np.float64)
    for b, beams in enumerate(res):
        for r, d in enumerate(beams):
            ln[b, r] = len(d.tokens)
            tok[b, r, :len(d.tokens)] = d.tokens
            sc[b, r] = d.score
    out[f"lg_{n}"], out[f"par_{n}"] = lg, np.array([W, blank], np.int32)
    out[f"tok_{n}"], out[f"len_{n}"], out[f"sc_{n}"] = tok, ln, sc
    out[f"lp_{n}"] = torch.log_softmax(torch.from_numpy(lg), dim=-1).numpy()
    n += 1
out["n_cases"] = np.array(n)
np.savez_compressed(os.path.join(HERE, "beam.npz"), **out)
print("wrote beam.npz:", n, "cases", os.path.getsize(os.path.join(HERE, "beam.npz")), "bytes")
