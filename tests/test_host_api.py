"""Host-side logic that needs no GPU: configuration, the parameter tree, the C-ABI library
(loads and exports every symbol include/vasr.h declares), failure without a CUDA device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def va():
    import velocity_asr
    return velocity_asr


def test_config_defaults_and_from_dict(va):
    c = va.VelocityASRConfig()
    assert (c.mel_bins, c.d_model, c.ssm_layers, c.ssm_state_dim, c.vocab_size, c.scan_mode) == \
        (80, 192, 8, 64, 1000, "parallel")
    c2 = va.VelocityASRConfig.from_dict({"d_model": 128, "unknown_key": 1, "scan_mode": "sequential"})
    assert c2.d_model == 128 and c2.scan_mode == "sequential"


def test_config_from_yaml(va, tmp_path):
    p = tmp_path / "model.yaml"
    p.write_text("input:\n  mel_bins: 80\n  n_fft: 400\nmodel:\n  d_model: 192\n  vocab_size: 500\n"
                 "ssm:\n  num_layers: 4\n  state_dim: 32\nglobal_context:\n  ssm_layers: 1\n"
                 "performance:\n  scan_mode: \"mamba\"\n")
    c = va.config_from_yaml(str(p))
    assert (c.vocab_size, c.ssm_layers, c.ssm_state_dim, c.global_ssm_layers, c.scan_mode) == (500, 4, 32, 1, "mamba")
    assert c.attention_dim == 48 and c.ssm_kernel_size == 4      # defaults for absent keys


def test_state_dict_layout(va):
    m = va.VELOCITYASR()
    sd = m.state_dict()
    assert len(sd) == 208 and m.count_parameters() == 6172696
    assert list(sd)[:4] == ["temporal_binding.conv.weight", "temporal_binding.conv.bias",
                            "temporal_binding.pos_encoding.pe_freq", "temporal_binding.pos_encoding.pe_time"]
    assert sd["local_ssm.layers.7.ssm.x_proj.weight"].shape == (128, 384)
    assert sd["global_context.global_ssm.layers.1.ssm.A_log"].shape == (32,)
    assert sd["global_context.fusion.gate_proj.0.weight"].shape == (192, 384)
    assert sd["ctc_head.proj.2.weight"].shape == (1000, 192)
    assert m.get_output_length(1501) == 751 and m.get_output_length(1000) == 500


def test_save_and_load_roundtrip(va, tmp_path):
    torch.manual_seed(3)
    m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential", vocab_size=300))
    path = str(tmp_path / "ckpt" / "model.pt")
    m.save_pretrained(path)
    m2 = va.from_pretrained(path)
    assert m2.config.scan_mode == "sequential" and m2.config.vocab_size == 300
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(NotImplementedError):
        va.from_pretrained("not-a-local-path")


def test_vocabulary_and_text(va):
    v = va.create_default_vocabulary(100)
    assert len(v) == 100 and v[:4] == ["<blank>", "<unk>", "<pad>", " "] and v[4] == "a" and v[99] == "<token_99>"
    d = va.CTCDecoder(v)
    assert d._tokens_to_text([4, 5, 3, 6, 1000]) == "ab c<unk>"
    assert d.text_to_tokens("ab c") == [4, 5, 3, 6]


def test_frontend_tables_match_reference_fixture(va, golden):
    from velocity_asr.audio import frontend_tables
    g = golden("frontend")
    fb, win = frontend_tables(80)
    assert (fb.numpy() == g["filterbank"]).all()       # same torch float32 ops -> same bits
    assert (win.numpy() == g["window"]).all()


def test_library_exports_every_declared_symbol():
    from velocity_asr import _native
    header = open(os.path.join(ROOT, "include", "vasr.h")).read()
    declared = set(re.findall(r"\b(vasr_[a-z0-9_]+)\s*\(", header))
    declared -= {"vasr_handle", "vasr_config"}
    assert os.path.exists(_native.LIB_PATH), "libvasr.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_native.exported_symbols())
    assert _native.lib().vasr_num_frames(240000) == 1501
    assert _native.lib().vasr_num_tokens(1501) == 751
    assert b"sm_100a" in _native.lib().vasr_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda(va):
    from velocity_asr import _native
    cfg = _native.VasrConfig(80, 192, 8, 64, 2, 4, 2, 32, 4, 48, 1000, 1)
    h = ctypes.c_void_p()
    rc = _native.lib().vasr_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == _native.ERR_CUDA and b"no CPU path" in _native.lib().vasr_last_error()
    # host tensors are staged to a CUDA device, never computed on the CPU: without a device every call raises
    with pytest.raises(RuntimeError, match="no CPU path"):
        va.VELOCITYASR()(torch.zeros(1, 10, 80))
    with pytest.raises(RuntimeError):
        va.compute_mel_spectrogram(torch.zeros(16000))
    with pytest.raises(RuntimeError):
        va.ctc_greedy_decode(torch.zeros(1, 5, 10))


def test_word_assembly_follows_the_reference_script(va):
    """scripts/transcribe.py:80-126 on a hand-made token stream: a word ends at the END frame of the separator
    that closes it, the last word at the end of the last token; leading / repeated separators add nothing."""
    vocab = va.create_default_vocabulary(100)
    t = {ch: vocab.index(ch) for ch in "hi yo"}
    tokens = [t[" "], t["h"], t["i"], t[" "], t[" "], t["y"], t["o"]]
    stamps = [(0, 1), (2, 3), (3, 5), (6, 8), (9, 10), (11, 12), (14, 17)]
    words = va.words_with_timestamps(tokens, stamps, vocab)
    sec = va.frames_to_seconds
    assert words == [{"word": "hi", "start": sec(2), "end": sec(8)}, {"word": "yo", "start": sec(11), "end": sec(17)}]
    assert sec(50) == 1.0                                  # 50 tokens = 100 mel frames = 1 s
    assert va.words_with_timestamps([], [], vocab) == []
    assert va.words_with_timestamps([1000], [(0, 1)], vocab)[0]["word"] == "<unk>"
