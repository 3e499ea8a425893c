"""The oracle (oracle/velocity_oracle.py) against fixtures produced by running the reference
(tests/golden/make_golden.py).  CPU only.  Tolerances: the fixtures are float32 torch output,
the oracle runs in float64, so differences are the reference's own float32 rounding."""
import numpy as np
import pytest
import torch

import fixtures_util as FU
import velocity_oracle as O


def rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def seeded_state_dict(amplified=False):
    """Same draw as the reference under torch.manual_seed(0), through the product's own
    parameter container (no reference import needed)."""
    import velocity_asr
    torch.manual_seed(FU.WEIGHT_SEED)
    sd = velocity_asr.VELOCITYASR(velocity_asr.VelocityASRConfig()).state_dict()
    if amplified:
        sd = FU.amplify_state_dict(sd)
    return sd


def test_filterbank_and_window(golden):
    g = golden("frontend")
    fb = O.mel_filterbank()
    assert fb.shape == (80, 201)
    assert np.abs(fb - g["filterbank"]).max() < 1e-5          # float32 pow/log10 differ in the last ulps
    assert ((fb > 0) == (g["filterbank"] > 0)).mean() > 0.999
    assert np.abs(O.hann_window(400, np.float32) - g["window"]).max() < 5e-7   # torch builds it in float32


def test_log_mel(golden):
    g = golden("frontend")
    audio = FU.synth_audio(2, 16000).numpy()
    fb = g["filterbank"]
    assert np.abs(O.log_mel(audio, filters=fb) - g["mel"]).max() < 1e-4
    assert np.abs(O.log_mel(audio, filters=fb, window=g["window"]) - g["mel"]).max() < 5e-5
    assert np.abs(O.log_mel(audio) - g["mel"]).max() < 1e-4
    assert np.abs(O.log_mel(audio, normalize=False, filters=fb) - g["mel_raw"]).max() < 1e-4
    odd = FU.synth_audio(3, 4037, seed=99).numpy()
    assert O.log_mel(odd).shape == g["mel_odd"].shape == (3, 26, 80)
    assert np.abs(O.log_mel(odd, filters=fb) - g["mel_odd"]).max() < 1e-4
    one = O.log_mel(audio[0, :800], filters=fb)
    assert one.shape == g["mel_1d"].shape == (6, 80)
    assert np.abs(one - g["mel_1d"]).max() < 1e-4


def test_weight_digest_matches_reference_init(golden):
    g = golden("model_small")
    sd = seeded_state_dict()
    assert len(sd) == 208
    np.testing.assert_allclose(FU.state_dict_digest(sd), g["digest"], rtol=0, atol=0)
    np.testing.assert_allclose(FU.state_dict_digest(FU.amplify_state_dict(sd)), g["digest_amp"], rtol=1e-12)


@pytest.mark.parametrize("mode", ["sequential", "parallel"])
@pytest.mark.parametrize("amp", [False, True])
def test_forward(golden, mode, amp):
    g = golden("model_small")
    mel = golden("frontend")["mel"]
    sd = {k: v.numpy() for k, v in seeded_state_dict(amp).items()}
    logits, f = O.forward(mel, sd, dict(scan_mode=mode), return_features=True)
    tag = f"{mode}{'_amp' if amp else ''}"
    assert rel(f["temporal_binding"], g[tag + "_tb"]) < 5e-6
    assert rel(f["local_features"], g[tag + "_local"]) < 2e-5
    assert rel(f["fused_features"], g[tag + "_fused"]) < 2e-5
    assert rel(logits, g[tag + "_logits"]) < 2e-5
    toks = O.ctc_greedy_decode(logits)
    ref = [[int(t) for t in row if t >= 0] for row in g[tag + "_tokens"]]
    assert toks == ref


def test_amplified_fixture_sees_the_ssm_branch(golden):
    """The two scan semantics must differ visibly at the logits of the amplified fixture,
    otherwise an end-to-end check cannot tell a wrong scan kernel from a right one."""
    g = golden("model_small")
    assert rel(g["sequential_amp_logits"], g["parallel_amp_logits"]) > 1e-2
    assert rel(g["sequential_logits"], g["parallel_logits"]) < 2e-3


def test_scan_ops(golden):
    g = golden("scan_ops")
    for case in g["cases"]:
        name, b, L, di, n, seed, st = eval(case)
        x, dt, A, Bm, Cm, D = [a.astype(np.float64) for a in FU.scan_inputs(b, L, di, n, seed, st)]
        assert rel(O.scan_sequential(x, dt, A, Bm, Cm, D), g[name + "_seq"]) < 1e-5
        assert rel(O.scan_parallel_streaming(x, dt, A, Bm, Cm, D), g[name + "_par"]) < 1e-5
        if L <= 40:
            assert rel(O.scan_parallel(x, dt, A, Bm, Cm, D), g[name + "_par"]) < 1e-5
        if L > 4:
            assert rel(g[name + "_seq"], g[name + "_par"]) > 0.05      # the quirk is not the recurrence


def test_blocks(golden):
    g = golden("blocks")
    sd = {k: v.numpy() for k, v in seeded_state_dict(True).items()}
    rs = np.random.RandomState(21)
    xin = rs.standard_normal((2, 75, 192)).astype(np.float32).astype(np.float64)
    xlong = rs.standard_normal((1, 1100, 192)).astype(np.float32).astype(np.float64)
    assert rel(O.ssm_block(xin, sd, "local_ssm.layers.0.", "sequential"), g["local_block0_seq"]) < 1e-5
    assert rel(O.ssm_block(xin, sd, "local_ssm.layers.0.", "parallel"), g["local_block0_par"]) < 1e-5
    assert rel(O.global_context(xin, sd, O.DEFAULT_CFG), g["global_context"]) < 1e-5
    assert O.pool_sizes(1100) == (137, 34) and O.pool_sizes(751) == (93, 23) and O.pool_sizes(40) == (40, 16)
    assert rel(O.global_context(xlong, sd, O.DEFAULT_CFG)[:, ::11], g["global_context_L1100"]) < 1e-5
    assert rel(O.ctc_head(xin, sd)[:, ::5], g["ctc_head"]) < 1e-5


def test_decode_known_answers(golden):
    g = golden("decode")
    for i in range(5):
        assert O.collapse_tokens(g[f"in{i}"]) == g[f"out{i}"].tolist()
        assert O.collapse_tokens(g[f"in{i}"], collapse_repeated=False) == g[f"out_nocollapse{i}"].tolist()
    assert O.collapse_tokens([0, 5, 5, 0, 5, 7, 7, 7, 0, 0, 3, 3, 5]) == [5, 5, 7, 3, 5]
    tie = np.zeros((1, 3, 6)); tie[0, 0, 2] = tie[0, 0, 4] = 1.0; tie[0, 1, 5] = 1.0
    assert O.ctc_greedy_decode(tie)[0] == g["tie_out"].tolist() == [2, 5]
    assert O.ctc_greedy_decode(np.zeros((2, 0, 5))) == [[], []]


def test_fake_quantize_arithmetic():
    x = np.array([[-1.0, 0.26, 0.5], [2.0, -0.1, 0.05]])
    s, zp = O.fake_quant_params(x, symmetric=True, per_channel=True)
    np.testing.assert_allclose(s[:, 0], [1 / 127, 2 / 127])
    q = O.fake_quantize(x, s, zp, symmetric=True)
    assert np.abs(q - x).max() <= s.max() / 2 + 1e-12
    s2, zp2 = O.fake_quant_params(x, symmetric=False, per_channel=False)
    np.testing.assert_allclose(s2, 3.0 / 255)
    assert np.abs(O.fake_quantize(x, s2, zp2, symmetric=False) - x).max() <= s2 / 2 + 1e-12
    assert O.fake_quantize(np.array([0.5, 1.5, 2.5]), 1.0, 0.0, True).tolist() == [0.0, 2.0, 2.0]  # half-to-even
