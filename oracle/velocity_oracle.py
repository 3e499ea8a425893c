"""CPU oracle for the VELOCITY-ASR v2 inference path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``velocity-asr_b200/`` imports this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker.

Parity status: PINNED.  The reference (shaderko/velocity-asr, pure PyTorch) holds
no golden vectors of its own, so every function here is
pinned by executing the reference itself in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``, checked by
``tests/test_oracle_golden.py``) and, when the reference tree is importable, by a
live comparison (``tests/test_oracle_vs_reference.py``).

Each function cites the reference file:line it restates.  All arithmetic is done
in ``dtype`` (float64 by default, so the oracle is the *tighter* side of every
tolerance; pass ``np.float32`` to imitate the reference's rounding).

State dicts are plain ``{name: np.ndarray}`` maps using the reference's
``state_dict()`` keys.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

try:  # scipy is in the image; erf is needed for the exact GELU
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

SAMPLE_RATE = 16000   # audio.py:15
N_FFT = 400           # audio.py:16
HOP_LENGTH = 160      # audio.py:17
N_MELS = 80           # audio.py:18
BLANK_TOKEN = 0       # decode.py:14


# --------------------------------------------------------------------------
# a1  log-mel front end  (audio.py:65-143, 146-199)
# --------------------------------------------------------------------------
def hann_window(n: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi i / n).  audio.py:97"""
    i = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * i / n)).astype(dtype)


def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace in float32: symmetric two-sided evaluation (ATen RangeFactories)."""
    start = np.float32(start)
    end = np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    lo = (start + step * i.astype(np.float32)).astype(np.float32)
    hi = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.float32)


def mel_filterbank(n_fft: int = N_FFT, n_mels: int = N_MELS,
                   sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """HTK-mel triangular filterbank (n_mels, n_fft//2+1), float32 like the reference.

    audio.py:164-199: freqs = linspace(0, sr/2, n_freqs); mel points equally spaced
    between hz_to_mel(0) and hz_to_mel(sr/2); row i = max(0, min(rising, falling))
    with a 1e-10 guard in both denominators.
    """
    f32 = np.float32
    n_freqs = n_fft // 2 + 1
    freqs = _linspace_f32(0.0, sample_rate / 2, n_freqs)
    mel_min = f32(2595) * np.log10(f32(1) + f32(0.0) / f32(700), dtype=f32)
    mel_max = f32(2595) * np.log10(f32(1) + f32(sample_rate / 2.0) / f32(700), dtype=f32)
    mel_pts = _linspace_f32(mel_min, mel_max, n_mels + 2)
    hz_pts = (f32(700) * (np.power(f32(10), mel_pts / f32(2595), dtype=f32) - f32(1))).astype(f32)
    lower = hz_pts[:-2, None]
    center = hz_pts[1:-1, None]
    upper = hz_pts[2:, None]
    rising = (freqs[None, :] - lower) / (center - lower + f32(1e-10))
    falling = (upper - freqs[None, :]) / (upper - center + f32(1e-10))
    return np.maximum(f32(0), np.minimum(rising, falling)).astype(f32)


def reflect_pad(audio: np.ndarray, pad: int) -> np.ndarray:
    """F.pad(mode='reflect') on the last axis (edge sample not repeated).  audio.py:100-101"""
    return np.pad(audio, [(0, 0)] * (audio.ndim - 1) + [(pad, pad)], mode="reflect")


def power_spectrogram(audio: np.ndarray, n_fft: int = N_FFT, hop: int = HOP_LENGTH,
                      dtype=np.float64, window: Optional[np.ndarray] = None) -> np.ndarray:
    """|STFT|^2, (B, n_fft//2+1, T): reflect pad n_fft//2, frames of n_fft every hop,
    periodic Hann, one-sided DFT, center=False.  audio.py:97-115"""
    audio = np.asarray(audio, dtype=dtype)
    x = reflect_pad(audio, n_fft // 2)
    n_frames = 1 + (x.shape[-1] - n_fft) // hop
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    win = hann_window(n_fft, dtype) if window is None else np.asarray(window, dtype=dtype)
    frames = x[:, idx] * win[None, None, :]
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)
    power = (spec.real ** 2 + spec.imag ** 2).astype(dtype)
    return np.transpose(power, (0, 2, 1))


def log_mel(audio: np.ndarray, sample_rate: int = SAMPLE_RATE, n_fft: int = N_FFT,
            hop_length: int = HOP_LENGTH, n_mels: int = N_MELS, normalize: bool = True,
            dtype=np.float64, filters: Optional[np.ndarray] = None,
            window: Optional[np.ndarray] = None) -> np.ndarray:
    """compute_mel_spectrogram: (S,)|(B,S) -> (T,n_mels)|(B,T,n_mels).  audio.py:65-143

    log(mel + 1e-10), then per (utterance, bin) (x - mean_T) / (std_T(unbiased) + 1e-10).
    """
    audio = np.asarray(audio)
    squeeze = audio.ndim == 1
    if squeeze:
        audio = audio[None]
    power = power_spectrogram(audio, n_fft, hop_length, dtype, window)
    fb = (mel_filterbank(n_fft, n_mels, sample_rate) if filters is None else filters).astype(dtype)
    mel = np.einsum("mf,bft->bmt", fb, power)
    mel = np.log(mel + dtype(1e-10))
    if normalize:
        mean = mel.mean(axis=-1, keepdims=True)
        std = mel.std(axis=-1, keepdims=True, ddof=1)
        mel = (mel - mean) / (std + dtype(1e-10))
    mel = np.transpose(mel, (0, 2, 1))
    return mel[0] if squeeze else mel


# --------------------------------------------------------------------------
# elementwise helpers (ATen semantics named in SURVEY.md section 8c)
# --------------------------------------------------------------------------
def layer_norm(x, weight, bias, eps: float = 1e-5):
    """F.layer_norm over the last axis, biased variance."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * weight + bias


def gelu(x):
    """nn.GELU() (approximate='none'): x * Phi(x) with the exact erf."""
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def softplus(x):
    """F.softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x))."""
    xs = np.minimum(x, 20.0)
    return np.where(x > 20.0, x, np.log1p(np.exp(xs)))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def silu(x):
    return x * sigmoid(x)


def linear(x, w, b=None):
    """F.linear: x @ w.T + b, w is (out, in)."""
    y = x @ w.T
    return y if b is None else y + b


class QuantState:
    """Config 5: which modules are QuantizedLinear / QuantizedConv1d (quantize.py:269-322 with
    ssm_state_fp32: everything outside sub-trees whose name contains "ssm"), the output FakeQuantize
    parameters gathered so far, and whether the nodes are in training mode (calibrating)."""
    MODULES = (
        "temporal_binding.conv",
        "global_context.pool1.pool_proj", "global_context.pool2.pool_proj",
        "global_context.cross_attention.q_proj", "global_context.cross_attention.k_proj",
        "global_context.cross_attention.v_proj", "global_context.cross_attention.out_proj",
        "global_context.fusion.gate_proj.0", "global_context.fusion.local_proj",
        "global_context.fusion.global_proj", "global_context.fusion.out_proj",
        "ctc_head.proj.2",
    )

    def __init__(self, act=None, calibrating: bool = False):
        self.act = dict(act or {})          # module -> (scale, zero_point)
        self.calibrating = calibrating


def quantized_weight(w):
    """QuantizedLinear / QuantizedConv1d weight path: per-output-channel symmetric int8 FakeQuantize,
    forward value x + (dq - x).  quantize.py:99-133, 182-183"""
    scale, zp = fake_quant_params(w, symmetric=True, per_channel=True)
    return w + (fake_quantize(w, scale, zp, symmetric=True) - w)


def quantized_output(out, name: str, q: "QuantState"):
    """The activation_quantizer of a quantised module (asymmetric per-tensor uint8): in training mode it
    first takes its scale / zero point from `out` (quantize.py:86-88); un-calibrated nodes pass through
    (quantize.py:82-84)."""
    if q.calibrating:
        scale, zp = fake_quant_params(out, symmetric=False, per_channel=False)
        q.act[name] = (float(scale), float(zp))
    if name not in q.act:
        return out
    scale, zp = q.act[name]
    return out + (fake_quantize(out, scale, zp, symmetric=False) - out)


def qlinear(x, sd, name: str, q, dtype=np.float64):
    """nn.Linear `name` of the state dict, or its QuantizedLinear replacement when q is given.
    quantize.py:180-191"""
    w = sd[name + ".weight"].astype(dtype)
    b = sd[name + ".bias"].astype(dtype)
    if q is None:
        return linear(x, w, b)
    return quantized_output(linear(x, quantized_weight(w), b), name, q)


# --------------------------------------------------------------------------
# a2  temporal binding  (model.py:94-127, 176-202)
# --------------------------------------------------------------------------
def positional_time_table(n_rows: int, d_model: int = 192, dtype=np.float64) -> np.ndarray:
    """pe_time rows [0, n_rows): interleaved sin/cos over d_model//2 channels.  model.py:94-100
    (the buffer in the reference has 5000 rows; config 4 regenerates it longer)."""
    half = d_model // 2
    pos = np.arange(n_rows, dtype=np.float32)[:, None]
    div = np.exp(np.arange(0, half, 2, dtype=np.float32) * np.float32(-math.log(10000.0) / half)).astype(np.float32)
    ang = (pos * div[None, :]).astype(np.float32)
    pe = np.zeros((n_rows, half), dtype=np.float32)
    pe[:, 0::2] = np.sin(ang)
    pe[:, 1::2] = np.cos(ang)
    return pe.astype(dtype)


def temporal_binding(mel: np.ndarray, sd: Dict[str, np.ndarray], dtype=np.float64, q=None) -> np.ndarray:
    """Conv1d(mel_bins->d_model, k=3, stride=2, pad=1) -> GELU -> + [pe_time | pe_freq] -> LN.

    model.py:187-200.  (B,T,80) -> (B,(T+1)//2,192)."""
    p = "temporal_binding."
    w = sd[p + "conv.weight"].astype(dtype)          # (192, 80, 3)
    if q is not None:
        w = quantized_weight(w)                      # QuantizedConv1d, quantize.py:248-266
    b = sd[p + "conv.bias"].astype(dtype)
    mel = np.asarray(mel, dtype=dtype)
    B, T, C = mel.shape
    L = (T + 2 - 3) // 2 + 1
    xp = np.zeros((B, T + 2, C), dtype=dtype)
    xp[:, 1:T + 1] = mel
    out = np.zeros((B, L, w.shape[0]), dtype=dtype)
    for j in range(3):
        out += xp[:, j:j + 2 * L:2][:, :L] @ w[:, :, j].T
    out = out + b
    if q is not None:
        out = quantized_output(out, "temporal_binding.conv", q)
    out = gelu(out)
    pe_time = sd[p + "pos_encoding.pe_time"].astype(dtype)
    if pe_time.shape[0] < L:
        raise RuntimeError(f"pe_time has {pe_time.shape[0]} rows < {L} tokens (model.py:125)")
    pe_freq = sd[p + "pos_encoding.pe_freq"].astype(dtype).reshape(1, 1, -1)
    pos = np.concatenate([np.broadcast_to(pe_time[None, :L], (1, L, pe_time.shape[1])),
                          np.broadcast_to(pe_freq, (1, L, pe_freq.shape[-1]))], axis=-1)
    out = out + pos
    return layer_norm(out, sd[p + "norm.weight"].astype(dtype), sd[p + "norm.bias"].astype(dtype))


# --------------------------------------------------------------------------
# a5  selective scan, three semantics  (ssm.py:134-337)
# --------------------------------------------------------------------------
def scan_sequential(x, dt, A, Bm, Cm, D=None):
    """True recurrence h = exp(dt*A) h + x*dt*B ; y = <h, C> (+ x*D).  ssm.py:146-171
    (== the 'mamba' branch, ssm.py:309-337).  x,dt:(B,L,Di) A:(N,) Bm,Cm:(B,L,N)."""
    Bsz, L, Di = x.shape
    h = np.zeros((Bsz, Di, A.shape[0]), dtype=x.dtype)
    y = np.empty_like(x)
    for t in range(L):
        dA = np.exp(dt[:, t, :, None] * A[None, None, :])
        dB = dt[:, t, :, None] * Bm[:, t, None, :]
        h = dA * h + x[:, t, :, None] * dB
        y[:, t] = np.einsum("bdn,bn->bd", h, Cm[:, t])
    return y if D is None else y + x * D


def _tree_scan_literal(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Literal restatement of _associative_scan (ssm.py:228-295) on (B,L,...) arrays:
    pad to 2^ceil(log2 L) with (1,0); correct up-sweep; root := identity; down-sweep whose
    right child becomes a_R*a_L , a_R(new)*b_L + b_R  (the mis-ordered combine, :283-284)."""
    L = a.shape[1]
    log_len = math.ceil(math.log2(max(L, 1)))
    P = 2 ** log_len
    if P > L:
        padw = [(0, 0), (0, P - L)] + [(0, 0)] * (a.ndim - 2)
        a = np.pad(a, padw, constant_values=1.0)
        b = np.pad(b, padw, constant_values=0.0)
    a = a.copy()
    b = b.copy()
    stride = 1
    for _ in range(log_len):
        r = np.arange(2 * stride - 1, P, 2 * stride)
        l = r - stride
        a_l, b_l, a_r, b_r = a[:, l], b[:, l], a[:, r], b[:, r]
        a[:, r] = a_r * a_l
        b[:, r] = a_r * b_l + b_r
        stride *= 2
    a[:, -1] = 1.0
    b[:, -1] = 0.0
    stride = P // 2
    for _ in range(log_len):
        r = np.arange(2 * stride - 1, P, 2 * stride)
        l = r - stride
        a_lo, b_lo = a[:, l].copy(), b[:, l].copy()
        a[:, l] = a[:, r]
        b[:, l] = b[:, r]
        a[:, r] = a[:, r] * a_lo
        b[:, r] = a[:, r] * b_lo + b[:, r]
        stride //= 2
    return b[:, :L]


def scan_parallel(x, dt, A, Bm, Cm, D=None):
    """scan_mode='parallel' exactly as shipped (ssm.py:193-214 + :228-295), quirk included.
    Materialises (B,L,Di,N); use small shapes."""
    dA = np.exp(dt[..., None] * A.reshape(1, 1, 1, -1))
    xdB = x[..., None] * (dt[..., None] * Bm[:, :, None, :])
    h = _tree_scan_literal(dA, xdB)
    y = np.einsum("bldn,bln->bld", h, Cm)
    return y if D is None else y + x * D


def scan_parallel_streaming(x, dt, A, Bm, Cm, D=None):
    """Same function as scan_parallel, evaluated as a left-to-right sweep (the form the
    CUDA kernel uses; O(log L) live state per (d,n)).  With H[e] the true state after e
    steps, S[e] = sum_{s<e} dt_s, parent(t) = t & (t-1):

        hP[0] = 0 ; hP[t] = hP[parent] + exp(A*S[t]) * (H[t] - exp(A*(S[t]-S[parent])) * H[parent])

    which is the closed form of SURVEY.md section 5.7 (aligned power-of-two blocks following
    the set bits of t, block k weighted by the decay of blocks 1..k).  y[t] = <hP[t], C[t]>."""
    Bsz, L, Di = x.shape
    N = A.shape[0]
    H = np.zeros((Bsz, Di, N), dtype=x.dtype)
    S = np.zeros((Bsz, Di), dtype=np.float64)
    nlev = max(1, L.bit_length() + 1)
    # anc[l] = (hP, H, S) of the most recent index whose trailing-zero count is >= l
    anc_hp = np.zeros((nlev, Bsz, Di, N), dtype=x.dtype)
    anc_H = np.zeros((nlev, Bsz, Di, N), dtype=x.dtype)
    anc_S = np.zeros((nlev, Bsz, Di), dtype=np.float64)
    y = np.empty_like(x)
    for t in range(L):
        if t == 0:
            hp = np.zeros_like(H)
        else:
            tz = (t & -t).bit_length() - 1
            php, pH, pS = anc_hp[tz + 1], anc_H[tz + 1], anc_S[tz + 1]
            q = np.exp(A[None, None, :] * S[..., None]).astype(x.dtype)
            pdec = np.exp(A[None, None, :] * (S - pS)[..., None]).astype(x.dtype)
            hp = php + q * (H - pdec * pH)
            for l in range(tz + 1):
                anc_hp[l], anc_H[l], anc_S[l] = hp, H, S
        y[:, t] = np.einsum("bdn,bn->bd", hp, Cm[:, t])
        # advance the true recurrence to H[t+1], S[t+1]
        dA = np.exp(dt[:, t, :, None] * A[None, None, :])
        H = dA * H + x[:, t, :, None] * (dt[:, t, :, None] * Bm[:, t, None, :])
        S = S + dt[:, t].astype(np.float64)
    return y if D is None else y + x * D


def selective_scan(x, dt, A, Bm, Cm, D=None, mode: str = "sequential"):
    """Dispatch on scan_mode.  ssm.py:119-126"""
    if mode in ("sequential", "mamba"):
        return scan_sequential(x, dt, A, Bm, Cm, D)
    if mode == "parallel":
        return scan_parallel_streaming(x, dt, A, Bm, Cm, D)
    raise ValueError(f"Unknown scan_mode: {mode}")


# --------------------------------------------------------------------------
# a3/a4/a6  SSM blocks  (ssm.py:92-132, 404-427, 491-505, 542-556)
# --------------------------------------------------------------------------
def causal_depthwise_conv(x, w, b):
    """Conv1d(groups=C, k, padding=k-1)[..., :L] on (B,L,C): out[t] = b + sum_j w[c,j] x[t+j-(k-1)].
    ssm.py:411-414.  w: (C,1,k)."""
    Bsz, L, C = x.shape
    k = w.shape[-1]
    xp = np.concatenate([np.zeros((Bsz, k - 1, C), dtype=x.dtype), x], axis=1)
    out = np.zeros_like(x)
    for j in range(k):
        out += xp[:, j:j + L] * w[:, 0, j][None, None, :]
    return out + b


def selective_ssm(x, sd, p, mode, dtype=np.float64):
    """SelectiveSSM.forward.  ssm.py:105-130"""
    g = lambda k: sd[p + k].astype(dtype)
    xz = linear(x, g("in_proj.weight"))
    di = xz.shape[-1] // 2
    xi, z = xz[..., :di], xz[..., di:]
    bc = linear(xi, g("x_proj.weight"))
    n = bc.shape[-1] // 2
    Bm, Cm = bc[..., :n], bc[..., n:]
    dt = softplus(linear(xi, g("dt_proj.weight"), g("dt_proj.bias")))
    A = -np.exp(g("A_log"))
    y = selective_scan(xi, dt, A, Bm, Cm, g("D"), mode)
    return linear(y * silu(z), g("out_proj.weight"))


def ssm_block(x, sd, p, mode, dtype=np.float64):
    """SSMBlock._forward_impl (dropout = identity in eval).  ssm.py:404-427"""
    g = lambda k: sd[p + k].astype(dtype)
    u = layer_norm(x, g("norm1.weight"), g("norm1.bias"))
    u = causal_depthwise_conv(u, g("conv.weight"), g("conv.bias"))
    x = x + selective_ssm(u, sd, p + "ssm.", mode, dtype)
    v = layer_norm(x, g("norm2.weight"), g("norm2.bias"))
    v = linear(gelu(linear(v, g("ffn.0.weight"), g("ffn.0.bias"))), g("ffn.3.weight"), g("ffn.3.bias"))
    return x + v


def ssm_stack(x, sd, p, n_layers, mode, dtype=np.float64):
    """LocalSSMProcessor / GlobalSSM: blocks then a final LayerNorm.  ssm.py:501-505, 552-556"""
    for i in range(n_layers):
        x = ssm_block(x, sd, f"{p}layers.{i}.", mode, dtype)
    return layer_norm(x, sd[p + "norm.weight"].astype(dtype), sd[p + "norm.bias"].astype(dtype))


# --------------------------------------------------------------------------
# a7-a10  hierarchical global context  (attention.py)
# --------------------------------------------------------------------------
def pool_sizes(L: int) -> Tuple[int, int]:
    """K1 = min(max(64, L//8), L); K2 = min(min(64, max(16, K1//4)), K1).  attention.py:37-44,67"""
    k1 = min(max(64, L // 8), L)
    k2 = min(min(64, max(16, k1 // 4)), k1)
    return k1, k2


def adaptive_avg_pool(x, K: int):
    """F.adaptive_avg_pool1d over time on (B,L,C): window i = [floor(iL/K), ceil((i+1)L/K)).
    attention.py:71-73"""
    Bsz, L, C = x.shape
    out = np.empty((Bsz, K, C), dtype=x.dtype)
    for i in range(K):
        s = (i * L) // K
        e = -((-(i + 1) * L) // K)
        out[:, i] = x[:, s:e].mean(axis=1)
    return out


def multi_head_attention(q_in, kv_in, sd, p, heads, dtype=np.float64, qs=None):
    """MultiHeadAttention.forward (no mask, dropout = identity).  attention.py:135-162"""
    g = lambda k: sd[p + k].astype(dtype)
    q = qlinear(q_in, sd, p + "q_proj", qs, dtype)
    k = qlinear(kv_in, sd, p + "k_proj", qs, dtype)
    v = qlinear(kv_in, sd, p + "v_proj", qs, dtype)
    Bsz, Lq, Adim = q.shape
    hd = Adim // heads
    q = q.reshape(Bsz, Lq, heads, hd).transpose(0, 2, 1, 3)
    k = k.reshape(Bsz, -1, heads, hd).transpose(0, 2, 1, 3)
    v = v.reshape(Bsz, -1, heads, hd).transpose(0, 2, 1, 3)
    s = q @ k.transpose(0, 1, 3, 2) / math.sqrt(hd)
    s = s - s.max(axis=-1, keepdims=True)
    a = np.exp(s)
    a = a / a.sum(axis=-1, keepdims=True)
    o = (a @ v).transpose(0, 2, 1, 3).reshape(Bsz, Lq, Adim)
    return qlinear(o, sd, p + "out_proj", qs, dtype)


def gated_fusion(loc, ctx, sd, p, dtype=np.float64, q=None):
    """g = sigmoid(W_g [loc|ctx]); out = W_o (g * W_l loc + (1-g) * W_c ctx).  attention.py:207-218"""
    g_ = lambda k: sd[p + k].astype(dtype)
    p_ = p.rstrip(".")
    gate = sigmoid(qlinear(np.concatenate([loc, ctx], axis=-1), sd, p_ + ".gate_proj.0", q, dtype))
    lt = qlinear(loc, sd, p_ + ".local_proj", q, dtype)
    gt = qlinear(ctx, sd, p_ + ".global_proj", q, dtype)
    return qlinear(gate * lt + (1.0 - gate) * gt, sd, p_ + ".out_proj", q, dtype)


def global_context(local, sd, cfg, dtype=np.float64, q=None):
    """HierarchicalGlobalContext.forward.  attention.py:296-319.  GlobalSSM blocks are built
    without a scan_mode argument (ssm.py:529-538) so they always run the 'parallel' scan."""
    p = "global_context."
    g = lambda k: sd[p + k].astype(dtype)
    L = local.shape[1]
    k1, k2 = pool_sizes(L)
    x1 = qlinear(adaptive_avg_pool(local, k1), sd, p + "pool1.pool_proj", q, dtype)
    xs = ssm_stack(x1, sd, p + "global_ssm.", cfg["global_ssm_layers"], "parallel", dtype)
    x2 = qlinear(adaptive_avg_pool(xs, k2), sd, p + "pool2.pool_proj", q, dtype)
    kv = layer_norm(x2, g("norm1.weight"), g("norm1.bias"))
    qn = layer_norm(local, g("norm2.weight"), g("norm2.bias"))
    ctx = multi_head_attention(qn, kv, sd, p + "cross_attention.", cfg["attention_heads"], dtype, qs=q)
    return gated_fusion(local, ctx, sd, p + "fusion.", dtype, q=q)


# --------------------------------------------------------------------------
# a11-a13  CTC head, whole forward, greedy decode
# --------------------------------------------------------------------------
def ctc_head(x, sd, dtype=np.float64, q=None):
    """LayerNorm -> Linear(d_model -> vocab).  model.py:223-227"""
    g = lambda k: sd["ctc_head.proj." + k].astype(dtype)
    return qlinear(layer_norm(x, g("0.weight"), g("0.bias")), sd, "ctc_head.proj.2", q, dtype)


DEFAULT_CFG = dict(mel_bins=80, d_model=192, ssm_layers=8, ssm_state_dim=64, ssm_expand_ratio=2,
                   ssm_kernel_size=4, global_ssm_layers=2, global_ssm_state_dim=32,
                   attention_heads=4, attention_dim=48, vocab_size=1000, scan_mode="parallel")


def forward(mel, sd, cfg=None, dtype=np.float64, return_features: bool = False, quant=None):
    """VELOCITYASR.forward.  model.py:350-368.  quant: a QuantState -> the model after
    prepare_model_for_qat with its weights kept (config 5)."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    x = temporal_binding(mel, sd, dtype, q=quant)
    local = ssm_stack(x, sd, "local_ssm.", cfg["ssm_layers"], cfg["scan_mode"], dtype)
    fused = global_context(local, sd, cfg, dtype, q=quant)
    logits = ctc_head(fused, sd, dtype, q=quant)
    if return_features:
        return logits, {"temporal_binding": x, "local_features": local, "fused_features": fused}
    return logits


def collapse_tokens(pred, blank_token: int = BLANK_TOKEN, collapse_repeated: bool = True) -> List[int]:
    """The per-utterance loop of ctc_greedy_decode.  decode.py:51-69 (a blank resets `prev`)."""
    out: List[int] = []
    prev = None
    for tok in pred:
        tok = int(tok)
        if tok == blank_token:
            prev = None
            continue
        if collapse_repeated and tok == prev:
            continue
        out.append(tok)
        prev = tok
    return out


def ctc_greedy_decode(logits, blank_token: int = BLANK_TOKEN, collapse_repeated: bool = True) -> List[List[int]]:
    """argmax over V (ties -> lowest index) then collapse.  decode.py:46-69"""
    pred = np.argmax(np.asarray(logits), axis=-1)
    return [collapse_tokens(row, blank_token, collapse_repeated) for row in pred]


def ctc_greedy_decode_with_timestamps(logits, blank_token: int = BLANK_TOKEN):
    """decode.py:74-125, loop restated: (tokens, [(start_frame, end_frame)]) per utterance."""
    out = []
    for pred in np.argmax(np.asarray(logits), axis=-1):
        tokens, stamps, prev, start = [], [], None, 0
        for i, tok in enumerate(pred):
            tok = int(tok)
            if tok == blank_token:
                if prev is not None and prev != blank_token:
                    stamps.append((start, i))
                prev = tok
                continue
            if tok != prev:
                if prev is not None and prev != blank_token:
                    stamps.append((start, i))
                tokens.append(tok)
                start = i
            prev = tok
        if prev is not None and prev != blank_token:
            stamps.append((start, len(pred)))
        out.append((tokens, stamps))
    return out


def log_softmax32(logits):
    """F.log_softmax in fp32 (decode.py:152): (x - max) - log(sum(exp(x - max)))."""
    x = np.asarray(logits, dtype=np.float32)
    sh = x - x.max(axis=-1, keepdims=True)
    return (sh - np.log(np.exp(sh).sum(axis=-1, keepdims=True, dtype=np.float32))).astype(np.float32)


def ctc_beam_search(logits, beam_width: int = 10, blank_token: int = BLANK_TOKEN, log_probs=None):
    """decode.py:128-217 (lm_scorer=None), restated.  Per utterance a list of (tokens, score), best first.
    One hypothesis per collapsed prefix; a frame offers every hypothesis the blank (prefix kept), its own last
    token again (prefix kept) or any other token (prefix grown); equal prefixes keep the better score, the
    earlier offer on a tie; the `beam_width` best survive, earlier offers first among equals.  Scores are
    Python floats (fp64) fed with fp32 log-probs.  `log_probs` overrides the fp32 log_softmax (used to hand
    the oracle the reference's own table)."""
    lp_all = log_softmax32(logits) if log_probs is None else np.asarray(log_probs)
    out = []
    for lp_utt in lp_all:
        hyps = [((), 0.0, None)]                       # (prefix, score, last token), ranked
        for lp in lp_utt:
            offers = {}                                # prefix -> [score, last]; dicts keep first-offer order
            def offer(prefix, sc, tok):
                cur = offers.get(prefix)
                if cur is None:
                    offers[prefix] = [sc, tok]
                elif cur[0] < sc:
                    cur[0], cur[1] = sc, tok
            for prefix, sc, last in hyps:
                offer(prefix, sc + float(lp[blank_token]), blank_token)
                for tok in range(lp.shape[0]):
                    if tok != blank_token:
                        offer(prefix if tok == last else prefix + (tok,), sc + float(lp[tok]), tok)
            ranked = sorted(offers.items(), key=lambda kv: -kv[1][0])[:beam_width]   # stable
            hyps = [(k, v[0], v[1]) for k, v in ranked]
        out.append([(list(k), sc) for k, sc, _ in hyps])
    return out


def transcribe(audio, sd, cfg=None, dtype=np.float64) -> List[List[int]]:
    """load->mel->model->greedy, the order of scripts/transcribe.py:69-82."""
    return ctc_greedy_decode(forward(log_mel(audio, dtype=dtype), sd, cfg, dtype))


# --------------------------------------------------------------------------
# a14  FakeQuantize arithmetic (config 5)  (quantize.py:99-133)
# --------------------------------------------------------------------------
def fake_quant_params(x, symmetric: bool, per_channel: bool, bits: int = 8):
    """calibrate(): symmetric -> scale = absmax/127, zp = 0; asymmetric -> scale = (max-min)/255,
    zp = -min/scale (not rounded); scale clamped >= 1e-10.  quantize.py:111-121"""
    axes = tuple(range(1, x.ndim)) if per_channel else None
    if symmetric:
        qmax = 2 ** (bits - 1) - 1
        scale = np.abs(x).max(axis=axes, keepdims=per_channel) / qmax
        zp = np.zeros_like(scale)
    else:
        lo = x.min(axis=axes, keepdims=per_channel)
        hi = x.max(axis=axes, keepdims=per_channel)
        scale = (hi - lo) / (2 ** bits - 1)
        scale = np.maximum(scale, 1e-10)
        zp = 0 - lo / scale
    return np.maximum(scale, 1e-10), zp


def fake_quantize(x, scale, zp, symmetric: bool, bits: int = 8):
    """q = clamp(round_half_even(x/scale + zp)); dequant (q - zp) * scale.  quantize.py:123-133"""
    if symmetric:
        qmin, qmax = -(2 ** (bits - 1)), 2 ** (bits - 1) - 1
    else:
        qmin, qmax = 0, 2 ** bits - 1
    q = np.clip(np.rint(x / scale + zp), qmin, qmax)
    return (q - zp) * scale
