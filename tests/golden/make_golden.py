"""Generate tests/golden/*.npz by RUNNING the unmodified reference (shaderko/velocity-asr).

Run in the build container, where /root/reference (or baseline/_ref) is importable:

    python tests/golden/make_golden.py

The reference ships no golden vectors, so these fixtures are the pin
for both the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py).
Inputs are regenerated from seeds on the test side (tests/fixtures_util.py); only outputs,
and a digest of the seeded weights, are stored.  torch CPU, float32, eval(), no_grad().
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refload import load_reference  # noqa: E402
import fixtures_util as FU  # noqa: E402

R = load_reference()
assert R is not None, "reference not importable"
torch.set_grad_enabled(False)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def build(mode, amplified=False):
    torch.manual_seed(FU.WEIGHT_SEED)
    m = R.VELOCITYASR(R.VelocityASRConfig(scan_mode=mode)).eval()
    if amplified:
        m.load_state_dict(FU.amplify_state_dict(m.state_dict()))
    return m


# ---- front end -------------------------------------------------------------
fb = R.audio._create_mel_filterbank(400, 80, 16000, torch.device("cpu")).numpy()
audio = FU.synth_audio(2, 16000)
mel = R.compute_mel_spectrogram(audio)
mel_raw = R.compute_mel_spectrogram(audio, normalize=False)
audio_odd = FU.synth_audio(3, 4000 + 37, seed=99)       # S not a multiple of hop
mel_odd = R.compute_mel_spectrogram(audio_odd)
mel_1d = R.compute_mel_spectrogram(audio[0, :800])       # shortest sensible clip, 1-D input
save("frontend", filterbank=fb, window=torch.hann_window(400).numpy(), mel=mel.numpy(),
     mel_raw=mel_raw.numpy(), mel_odd=mel_odd.numpy(), mel_1d=mel_1d.numpy())

# ---- whole model, both scan semantics, plain and amplified weights -----------
model_out = {}
for mode in ("sequential", "parallel"):
    for amp in (False, True):
        m = build(mode, amp)
        logits, f = m(mel, return_features=True)
        tag = f"{mode}{'_amp' if amp else ''}"
        model_out[tag + "_logits"] = logits.numpy()
        model_out[tag + "_tb"] = f["temporal_binding"].numpy()
        model_out[tag + "_local"] = f["local_features"].numpy()
        model_out[tag + "_fused"] = f["fused_features"].numpy()
        model_out[tag + "_tokens"] = np.array(
            [t + [-1] * (logits.shape[1] - len(t)) for t in R.ctc_greedy_decode(logits)], dtype=np.int32)
        if not amp and mode == "sequential":
            model_out["digest"] = FU.state_dict_digest(m.state_dict())
            model_out["digest_amp"] = FU.state_dict_digest(FU.amplify_state_dict(m.state_dict()))
            model_out["n_params"] = np.array(m.count_parameters())
save("model_small", **model_out)

# ---- BASELINE config 1: 1 x 10 s, end to end ---------------------------------
cfg1 = {}
audio1 = FU.synth_audio(1, 160000)
mel1 = R.compute_mel_spectrogram(audio1)
for mode in ("sequential", "parallel"):
    m = build(mode)
    lg = m(mel1)
    cfg1[mode + "_logits_sub"] = lg[:, ::25].numpy()
    cfg1[mode + "_argmax"] = lg.argmax(-1).numpy().astype(np.int32)
    toks = R.ctc_greedy_decode(lg)[0]
    cfg1[mode + "_tokens"] = np.array(toks, dtype=np.int32)
    # margin between best and second-best logit per frame (tie-freeness of the argmax)
    top2 = lg.topk(2, dim=-1).values
    cfg1[mode + "_margin"] = (top2[..., 0] - top2[..., 1]).numpy()
cfg1["mel_sub"] = mel1[:, ::50].numpy()
save("config1", **cfg1)

# ---- per-op selective scan with O(1) inputs ----------------------------------
scan = {}
cases = [("n64_L37", 2, 37, 384, 64, 11, True), ("n64_L100_gen", 2, 100, 384, 64, 12, False),
         ("n32_L93", 2, 93, 384, 32, 13, True), ("n64_L1", 1, 1, 384, 64, 14, True),
         ("n64_L130", 1, 130, 384, 64, 15, True)]
for name, b, L, di, n, seed, st in cases:
    x, dt, A, Bm, Cm, D = FU.scan_inputs(b, L, di, n, seed, st)
    ssm = R.SelectiveSSM(d_model=di // 2, state_dim=n, expand_ratio=2)
    ssm.D.data = torch.from_numpy(D)
    t = torch.from_numpy
    scan[name + "_seq"] = ssm._sequential_scan(t(x), t(dt), t(A), t(Bm), t(Cm)).numpy()
    scan[name + "_par"] = ssm._parallel_scan(t(x), t(dt), t(A), t(Bm), t(Cm)).numpy()
scan["cases"] = np.array([repr(c) for c in cases])
save("scan_ops", **scan)

# ---- one SSM block and the global-context module on O(1) activations ---------
blk = {}
m = build("sequential", amplified=True)
rs = np.random.RandomState(21)
xin = torch.from_numpy(rs.standard_normal((2, 75, 192)).astype(np.float32))
blk["local_block0_seq"] = m.local_ssm.layers[0](xin).numpy()
m_par = build("parallel", amplified=True)
blk["local_block0_par"] = m_par.local_ssm.layers[0](xin).numpy()
blk["global_context"] = m.global_context(xin).numpy()
xlong = torch.from_numpy(rs.standard_normal((1, 1100, 192)).astype(np.float32))
blk["global_context_L1100"] = m.global_context(xlong)[:, ::11].numpy()   # K1=137, K2=34
blk["ctc_head"] = m.ctc_head(xin).numpy()[:, ::5]
save("blocks", **blk)

# ---- greedy decode known answers (decode.py:46-69) ----------------------------
seqs = [[0, 5, 5, 0, 5, 7, 7, 7, 0, 0, 3, 3, 5], [0, 0, 0, 0], [4], [9, 9, 9], [1, 0, 1, 0, 1, 1, 2, 2, 0]]
dec = {}
for i, s in enumerate(seqs):
    lg = torch.nn.functional.one_hot(torch.tensor([s]), 10).float()
    dec[f"in{i}"] = np.array(s, dtype=np.int32)
    dec[f"out{i}"] = np.array(R.ctc_greedy_decode(lg)[0], dtype=np.int32)
    dec[f"out_nocollapse{i}"] = np.array(R.ctc_greedy_decode(lg, collapse_repeated=False)[0], dtype=np.int32)
tie = torch.zeros(1, 3, 6); tie[0, 0, 2] = tie[0, 0, 4] = 1.0; tie[0, 1, 5] = 1.0   # ties -> lowest index
dec["tie_out"] = np.array(R.ctc_greedy_decode(tie)[0], dtype=np.int32)
save("decode", **dec)
