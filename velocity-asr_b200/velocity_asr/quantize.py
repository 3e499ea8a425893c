"""Config 5 — the reference's fake-quantised model (velocity_asr/quantize.py) on the CUDA path.

Same entry points as the reference for the inference side: ``QuantizationConfig``,
``prepare_model_for_qat`` and ``calibrate_model``.  What they do here:

* ``prepare_model_for_qat(model)`` marks the 12 modules the reference replaces when
  ``ssm_state_fp32`` is set (quantize.py:291-293: every Linear / Conv1d outside sub-trees whose name
  contains "ssm") as quantised: per-output-channel symmetric int8 FakeQuantize of the weight, fp32
  matmul, per-tensor asymmetric uint8 FakeQuantize of the output (quantize.py:180-191, 248-266).
  Unlike the reference, which builds fresh ``nn.Linear`` / ``nn.Conv1d`` modules and thereby drops
  the trained weights (quantize.py:295-313), the model keeps its weights and its
  ``state_dict`` keys.
* ``calibrate_model(model, batches)`` runs each calibration batch through the model with the output
  FakeQuantize nodes in training mode (quantize.py:86-88): every node takes min / max of the tensor
  it is about to quantise, derives scale / zero point (quantize.py:99-121) and quantises with them;
  the values of the last batch stay (the reference's ``_update_scale_zp`` overwrites per call).  The
  reference's own ``calibrate_model`` never gathers statistics (eval-mode FakeQuantize returns its
  input, quantize.py:82-84) and ends with scale 1 / zero point 0, i.e. all-zero logits.
* Un-calibrated output nodes pass through, as in the reference (quantize.py:82-84).

The ONNX / onnxruntime export helpers of the reference are out of scope.
"""
import ctypes
from dataclasses import dataclass
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import _native

QUANTIZED_MODULES = (
    "temporal_binding.conv",
    "global_context.pool1.pool_proj", "global_context.pool2.pool_proj",
    "global_context.cross_attention.q_proj", "global_context.cross_attention.k_proj",
    "global_context.cross_attention.v_proj", "global_context.cross_attention.out_proj",
    "global_context.fusion.gate_proj.0", "global_context.fusion.local_proj",
    "global_context.fusion.global_proj", "global_context.fusion.out_proj",
    "ctc_head.proj.2",
)


@dataclass
class QuantizationConfig:
    """quantize.py:18-37"""
    weight_bits: int = 8
    activation_bits: int = 8
    per_channel_weights: bool = True
    ssm_state_fp32: bool = True
    num_calibration_batches: int = 100
    symmetric_weights: bool = True
    symmetric_activations: bool = False


def prepare_model_for_qat(model, config: Optional[QuantizationConfig] = None):
    """quantize.py:269-322 (inference side).  Returns the same model object, marked quantised."""
    config = config or QuantizationConfig()
    if (config.weight_bits, config.activation_bits, config.per_channel_weights, config.ssm_state_fp32,
            config.symmetric_weights, config.symmetric_activations) != (8, 8, True, True, True, False):
        raise NotImplementedError("only the reference's default QuantizationConfig is built for the CUDA path")
    model._quantized = True
    model._act_qparams = {}
    return model


def calibrate_model(model, calibration_dataloader: Iterable, num_batches: int = 100, device: str = "cuda"):
    """quantize.py:325-371.  Batches are mel tensors, (mel, ...) tuples or {'mel_spectrogram': mel} dicts."""
    if not getattr(model, "_quantized", False):
        raise RuntimeError("call prepare_model_for_qat(model) first")
    model.to(device).eval()
    dev = model._device()
    eng = model._engine(dev)
    seen = 0
    for batch_idx, batch in enumerate(calibration_dataloader):
        if batch_idx >= num_batches:
            break
        mel = batch["mel_spectrogram"] if isinstance(batch, dict) else batch[0] if isinstance(batch, (tuple, list)) else batch
        mel = mel.to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            _native.check(eng.lib.vasr_calibrate(eng.handle, _native.ptr(mel), mel.size(0), mel.size(1),
                                                 ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        seen += 1
    if seen:
        model._act_qparams = read_quant_params(model)
    return model


def read_quant_params(model) -> Dict[str, Tuple[float, float]]:
    """{module name: (scale, zero_point)} of the output FakeQuantize nodes, as now on the device."""
    eng = model._engine(model._device())
    out = {}
    for name in QUANTIZED_MODULES:
        s, z = ctypes.c_float(), ctypes.c_float()
        _native.check(eng.lib.vasr_get_quant_params(eng.handle, name.encode(), ctypes.byref(s), ctypes.byref(z)))
        out[name] = (s.value, z.value)
    return out


def run_quantized_module(model, module: str, x: torch.Tensor) -> torch.Tensor:
    """QuantizedLinear.forward / QuantizedConv1d.forward of one of the 12 replaced modules on its own
    (quantize.py:180-191, 248-266): parity seam.  x is the input of the PROJECTION that hosts the module: the
    module's own input, except that k_proj / v_proj share one projection and gate_proj.0 / local_proj / global_proj
    another, whose input is [local | ctx] (B, L, 2 d_model); for temporal_binding.conv x is the mel (B, T, mel_bins).
    Returns the module's output (B, rows, out_features)."""
    if not getattr(model, "_quantized", False):
        raise RuntimeError("call prepare_model_for_qat(model) first")
    dev = model._exec_device(x)
    eng = model._engine(dev)
    x = x.to(dev, torch.float32).contiguous()
    n, c0, nc = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    name = module.encode()
    _native.check(eng.lib.vasr_quant_site(eng.handle, name, None, 0, 0, None, ctypes.byref(n), ctypes.byref(c0),
                                          ctypes.byref(nc), None))
    B, R = x.size(0), x.size(1)
    rows = (R + 1) // 2 if module == "temporal_binding.conv" else R
    out = torch.empty(B, rows, n.value, device=dev, dtype=torch.float32)
    _native.check(eng.lib.vasr_quant_site(eng.handle, name, _native.ptr(x), B, R, _native.ptr(out), None, None, None,
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out[:, :, c0.value:c0.value + nc.value]
