"""Generate tests/golden/quant.npz (config 5) by RUNNING the unmodified reference.

    python tests/golden/make_golden_quant.py          (build container: /root/reference importable)

Recipe (SURVEY.md 5.8 — the reference's own calibrate_model is degenerate: it never gathers statistics
and ends with scale 1 / zero point 0):
  1. fp32 model, torch.manual_seed(0), scan_mode "sequential";
  2. velocity_asr.quantize.prepare_model_for_qat(model)   (replaces the 12 non-SSM Linear / Conv1d modules
     by QuantizedLinear / QuantizedConv1d — with fresh weights);
  3. copy the original weights and biases back into the replacements' .linear / .conv;
  4. calibration: ONE forward on the calibration batch with every FakeQuantize node in training mode
     (quantize.py:86-88: each node updates its scale / zero point from the tensor it sees, then quantises)
     while everything else stays in eval mode (dropout off); then calibrated <- True, eval;
  5. forward on a second batch with the frozen parameters.
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
Q = sys.modules[R.__name__ + ".quantize"] if (R.__name__ + ".quantize") in sys.modules else __import__(
    R.__name__ + ".quantize", fromlist=["x"])
torch.set_grad_enabled(False)

torch.manual_seed(FU.WEIGHT_SEED)
model = R.VELOCITYASR(R.VelocityASRConfig(scan_mode="sequential")).eval()
sd0 = {k: v.clone() for k, v in model.state_dict().items()}
mel_cal = R.compute_mel_spectrogram(FU.synth_audio(2, 16000))
mel_test = R.compute_mel_spectrogram(FU.synth_audio(2, 16000, seed=77))
logits_fp32 = model(mel_test).numpy()

qmodel = Q.prepare_model_for_qat(model)
names = []
for name, mod in qmodel.named_modules():
    if isinstance(mod, (Q.QuantizedLinear, Q.QuantizedConv1d)):
        inner = mod.linear if isinstance(mod, Q.QuantizedLinear) else mod.conv
        inner.weight.copy_(sd0[name + ".weight"])
        inner.bias.copy_(sd0[name + ".bias"])
        names.append(name)
assert len(names) == 12, names
qmodel.eval()
fqs = [m for m in qmodel.modules() if isinstance(m, Q.FakeQuantize)]
for m in fqs:
    m.train()
logits_cal = qmodel(mel_cal).numpy()
for m in fqs:
    m.calibrated.fill_(True)
    m.eval()
mods = dict(qmodel.named_modules())
# input and output of each of the 12 quantised modules during the frozen forward: the per-module parity fixture
captured = {}
hooks = [mods[n].register_forward_hook(
    lambda m, inp, out, n=n: captured.__setitem__(n, (inp[0].detach().clone(), out.detach().clone()))) for n in names]
logits_q = qmodel(mel_test).numpy()
for hk in hooks:
    hk.remove()
per_module = {}
for n in names:
    xin, yout = captured[n]
    if n == "temporal_binding.conv":            # Conv1d sees (B, C, T) and returns (B, C, L): store (B, T, C) / (B, L, C)
        xin, yout = xin.transpose(1, 2), yout.transpose(1, 2)
    per_module[n + ":in"] = xin.contiguous().numpy()
    per_module[n + ":out"] = yout.contiguous().numpy()
mpath = os.path.join(HERE, "quant_modules.npz")
np.savez_compressed(mpath, names=np.array(names), **per_module)
print("quant_modules.npz", os.path.getsize(mpath) // 1024, "KiB")
act_scale = np.array([float(mods[n].activation_quantizer.scale) for n in names], dtype=np.float64)
act_zp = np.array([float(mods[n].activation_quantizer.zero_point) for n in names], dtype=np.float64)
# quantised weight of one module, as the forward pass sees it
wq_ctc = mods["ctc_head.proj.2"].weight_quantizer(mods["ctc_head.proj.2"].linear.weight).numpy()
path = os.path.join(HERE, "quant.npz")
np.savez_compressed(path, names=np.array(names), act_scale=act_scale, act_zp=act_zp, logits_cal=logits_cal,
                    logits_q=logits_q, logits_fp32=logits_fp32, wq_ctc_rows=wq_ctc[:8],
                    weights_digest=FU.state_dict_digest(sd0))
print("quant.npz", os.path.getsize(path) // 1024, "KiB;", names)
print("act scales", act_scale)
print("|logits_q - logits_fp32| max", np.abs(logits_q - logits_fp32).max(), "of", np.abs(logits_fp32).max())
