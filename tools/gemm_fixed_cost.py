"""Fixed cost of a projection launch: time at 1, 2, 3 full rounds of tiles (74 CTA pairs x 256 rows per round)
for the out_proj shape; the intercept of the line is what a launch costs beyond its tiles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
import torch
import velocity_asr as va
K, N, REPS = 384, 192, 8
g = torch.Generator(device="cuda").manual_seed(1)
w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
b = torch.randn(N, device="cuda", generator=g)
ws = va.split_tf32(w)
res = {}
for rounds in (1, 2, 3, 4):
    M = 74 * 256 * rounds
    xs = [torch.randn(M, K, device="cuda", generator=g) for _ in range(REPS)]
    rs = [torch.randn(M, N, device="cuda", generator=g) for _ in range(REPS)]
    outs = [torch.empty(M, N, device="cuda") for _ in range(REPS)]
    def go():
        for i in range(REPS):
            va.linear(xs[i], w, b, tensor_cores=True, weight_split=ws, residual=rs[i], out=outs[i])
    go(); go()
    ms = []
    for _ in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1) / REPS * 1e3)
    ms.sort(); res[rounds] = ms[len(ms) // 2]
    print(f"rounds {rounds}: M = {M}: {res[rounds]:.1f} us per launch")
per = (res[4] - res[1]) / 3
print(f"per round {per:.1f} us (ideal 12 k-blocks x 1152 clocks = 7.2 us at 1.92 GHz); fixed cost per launch {res[1] - per:.1f} us")
