"""Log-mel front end: compute_mel_spectrogram of velocity_asr/audio.py:65-143 on the GPU."""
import ctypes
from functools import lru_cache
from typing import Tuple

import torch

from . import _native

SAMPLE_RATE = 16000   # audio.py:15-18
N_FFT = 400
HOP_LENGTH = 160
N_MELS = 80


@lru_cache(maxsize=4)
def frontend_tables(n_mels: int = N_MELS) -> Tuple[torch.Tensor, torch.Tensor]:
    """(filterbank (n_mels, 201), window (400)) as float32 CPU tensors, evaluated with the same
    torch float32 ops as audio.py:97,164-199 — all rows at once instead of the reference's
    per-row loop, and once per process instead of on every call — so the values are
    bit-identical to the reference's on the same torch build."""
    n_freqs = N_FFT // 2 + 1
    freqs = torch.linspace(0, SAMPLE_RATE / 2, n_freqs)
    to_mel = lambda hz: 2595 * torch.log10(1 + hz / 700)
    mel_pts = torch.linspace(to_mel(torch.tensor(0.0)), to_mel(torch.tensor(SAMPLE_RATE / 2.0)), n_mels + 2)
    hz = 700 * (10 ** (mel_pts / 2595) - 1)
    lower, center, upper = hz[:-2, None], hz[1:-1, None], hz[2:, None]
    rising = (freqs[None, :] - lower) / (center - lower + 1e-10)
    falling = (upper - freqs[None, :]) / (upper - center + 1e-10)
    fb = torch.maximum(torch.zeros(1), torch.minimum(rising, falling)).contiguous()
    return fb, torch.hann_window(N_FFT)


def compute_mel_spectrogram(audio: torch.Tensor, sample_rate: int = SAMPLE_RATE, n_fft: int = N_FFT,
                            hop_length: int = HOP_LENGTH, n_mels: int = N_MELS,
                            normalize: bool = True) -> torch.Tensor:
    """audio (S,) | (B, S) -> (T, n_mels) | (B, T, n_mels), T = 1 + S // hop, on the device of `audio`.

    Same signature and result as audio.py:65-143.  The kernels are built for the model's
    front end (16 kHz, 400-point DFT, hop 160); other values raise NotImplementedError.
    The arithmetic always runs in the CUDA kernels: a CPU tensor (what scripts/transcribe.py:69 passes) is
    copied to the current CUDA device, the result is copied back — a host copy around the same kernels, not
    a CPU implementation; without a CUDA device the call raises."""
    if (sample_rate, n_fft, hop_length) != (SAMPLE_RATE, N_FFT, HOP_LENGTH):
        raise NotImplementedError("libvasr front end is fixed at sample_rate=16000, n_fft=400, hop_length=160")
    home = audio.device
    if home.type != "cuda":
        audio = audio.to(_native.default_cuda_device())
    squeeze = audio.dim() == 1
    if squeeze:
        audio = audio.unsqueeze(0)
    pcm = audio.to(torch.float32).contiguous()
    B, S = pcm.shape
    eng = _mel_engine(pcm.device, n_mels)
    T = 1 + S // HOP_LENGTH
    mel = torch.empty(B, T, n_mels, device=pcm.device, dtype=torch.float32)
    if B > 0:
        _native.check(eng.lib.vasr_log_mel(eng.handle, _native.ptr(pcm), B, S, int(bool(normalize)),
                                           _native.ptr(mel),
                                           ctypes.c_void_p(torch.cuda.current_stream(pcm.device).cuda_stream)))
    mel = mel.squeeze(0) if squeeze else mel
    return mel if home.type == "cuda" else mel.to(home)


_MEL_ENGINES = {}


def _mel_engine(device: torch.device, n_mels: int):
    """A weight-less handle that only carries the front-end tables."""
    from .config import VelocityASRConfig
    from .model import _Engine
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, n_mels)
    eng = _MEL_ENGINES.get(key)
    if eng is None:
        cfg = VelocityASRConfig(mel_bins=n_mels, ssm_layers=0, global_ssm_layers=0)
        eng = _Engine(cfg, torch.device("cuda", idx))
        fb, win = frontend_tables(n_mels)
        lib = eng.lib
        for k, v in (("frontend.mel_filterbank", fb), ("frontend.window", win)):
            _native.check(lib.vasr_set_weight(eng.handle, k.encode(), _native.ptr(v), v.numel()))
        _MEL_ENGINES[key] = eng
    return eng
