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


def _beam_arrays(res, W, L):
    tok = np.full((len(res), W, L), -1, np.int32)
    ln = np.full((len(res), W), -1, np.int32)
    sc = np.full((len(res), W), -np.inf)
    for b, beams in enumerate(res):
        for r, (t, s) in enumerate(beams):
            ln[b, r], sc[b, r] = len(t), s
            tok[b, r, :len(t)] = t
    return tok, ln, sc


def test_beam_search_matches_reference(golden):
    """ctc_beam_search (decode.py:128-217): with the reference's own log-prob table the restatement is
    bit-exact (prefixes, ranking, fp64 scores); with its own fp32 log_softmax the prefixes and ranking are
    the same and the scores agree to fp32 rounding of the table."""
    g = golden("beam")
    for i in range(int(g["n_cases"])):
        lg, (W, blank) = g[f"lg_{i}"], g[f"par_{i}"].tolist()
        tok, ln, sc = _beam_arrays(O.ctc_beam_search(lg, W, blank, log_probs=g[f"lp_{i}"]), W, lg.shape[1])
        assert np.array_equal(tok, g[f"tok_{i}"]) and np.array_equal(ln, g[f"len_{i}"]), i
        assert np.array_equal(sc, g[f"sc_{i}"]), i
        tok, ln, sc = _beam_arrays(O.ctc_beam_search(lg, W, blank), W, lg.shape[1])
        assert np.array_equal(tok, g[f"tok_{i}"]) and np.array_equal(ln, g[f"len_{i}"]), i
        assert np.allclose(sc, g[f"sc_{i}"], rtol=0, atol=1e-4), i
    assert O.ctc_beam_search(np.zeros((2, 0, 5), np.float32), 4) == [[([], 0.0)], [([], 0.0)]]


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


def test_quantized_forward_matches_reference(golden):
    """Config 5: the oracle's quantised forward (weights kept, output nodes calibrated by one forward in
    training mode) against the reference run by tests/golden/make_golden_quant.py.  Quantisation puts
    values on a grid, so a fp64-vs-fp32 rounding difference right at a grid boundary moves an element by
    one step: scales must agree tightly, logits almost everywhere, and never by more than a few steps."""
    import torch
    import velocity_asr as va
    g = golden("quant")
    torch.manual_seed(FU.WEIGHT_SEED)
    sd_t = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential")).state_dict()
    assert np.allclose(FU.state_dict_digest(sd_t), g["weights_digest"], rtol=0, atol=0)
    sd = {k: v.numpy() for k, v in sd_t.items()}
    mel_cal = O.log_mel(FU.synth_audio(2, 16000).numpy())
    mel_test = O.log_mel(FU.synth_audio(2, 16000, seed=77).numpy())
    cfg = dict(scan_mode="sequential")
    q = O.QuantState(calibrating=True)
    lc = O.forward(mel_cal, sd, cfg, quant=q)
    names = [str(n) for n in g["names"]]
    assert sorted(names) == sorted(O.QuantState.MODULES) == sorted(q.act)
    for n, s_ref, z_ref in zip(names, g["act_scale"], g["act_zp"]):
        s, z = q.act[n]
        # extremes downstream of a quantised tensor move with single grid-boundary flips upstream
        assert abs(s - s_ref) < 1e-2 * s_ref and abs(z - z_ref) < 1.0, (n, s, s_ref, z, z_ref)
    step = float(g["act_scale"][names.index("ctc_head.proj.2")])
    # own calibration: the output grid is shifted by the (tiny) differences in scale / zero point
    assert np.abs(lc - g["logits_cal"]).max() <= 4.0 * step and np.abs(lc - g["logits_cal"]).mean() < step
    # the reference's calibrated parameters: same grids, so equal up to boundary flips
    q = O.QuantState(act={n: (float(s_), float(z_)) for n, s_, z_ in zip(names, g["act_scale"], g["act_zp"])})
    lq = O.forward(mel_test, sd, cfg, quant=q)
    d = np.abs(lq - g["logits_q"])
    # eleven quantisers sit upstream of the logits: a value within rounding noise of a grid boundary flips
    # by one step and drags some downstream values across theirs.  Most logits are identical, the rest are
    # one (rarely two) steps of the last quantiser away, and the frame argmax is stable.
    assert (d < 1e-3).mean() > 0.8, (d < 1e-3).mean()
    assert d.max() <= 2.5 * step, (d.max(), step)
    assert (lq.argmax(-1) == g["logits_q"].argmax(-1)).mean() >= 0.95
    assert np.abs(g["logits_q"] - g["logits_fp32"]).max() > 5 * step       # quantisation is visible
    wq = O.quantized_weight(sd["ctc_head.proj.2.weight"].astype(np.float64))[:8]
    assert np.abs(wq - g["wq_ctc_rows"]).max() < 1e-7


@pytest.mark.parametrize("mode", ["sequential", "parallel"])
def test_config2_utterance_at_benchmark_size(golden, mode):
    """BASELINE configs[1] size: utterance 31 of the seed-1234 64 x 15 s batch (751 tokens) through the oracle,
    against the reference's own output at that size (tests/golden/make_golden_big.py), both scan semantics."""
    g = golden("config2")
    row = g["utts"].tolist().index(31)
    audio = FU.synth_audio(64, 240000)[31:32].numpy()
    mel = O.log_mel(audio)
    assert np.abs(mel[:, ::50] - g["mel_sub"][row:row + 1]).max() < 1e-4
    sd = {k: v.numpy() for k, v in seeded_state_dict().items()}
    logits = O.forward(mel, sd, dict(scan_mode=mode))
    assert logits.shape == (1, 751, 1000)
    assert rel(logits[:, ::25], g[mode + "_logits_sub"][row:row + 1]) < 1e-3
    safe = g[mode + "_margin"][row] > 1e-3
    assert (logits.argmax(-1)[0] == g[mode + "_argmax"][row])[safe].all()
    if safe.all():
        want = [t for t in g[mode + "_tokens"][row].tolist() if t >= 0]
        assert O.ctc_greedy_decode(logits)[0] == want


def test_quantized_modules_one_by_one(golden):
    """The oracle's QuantizedLinear / QuantizedConv1d arithmetic (quantize.py:180-191, 248-266) on the recorded input
    of each of the 12 modules of the reference's frozen forward: same grid point, one step at most at boundary ties."""
    g, gm = golden("quant"), golden("quant_modules")
    sd = {k: v.numpy() for k, v in seeded_state_dict().items()}
    for n, scale, zp in zip([str(x) for x in g["names"]], g["act_scale"], g["act_zp"]):
        x, want = gm[n + ":in"].astype(np.float32), gm[n + ":out"]
        w = O.quantized_weight(sd[n + ".weight"].astype(np.float32))
        if n == "temporal_binding.conv":                       # stride 2, kernel 3, padding 1 as a projection of 3 frames
            xp = np.pad(x, ((0, 0), (1, 1), (0, 0)))
            L = want.shape[1]
            frames = np.stack([xp[:, 2 * l:2 * l + 3] for l in range(L)], 1)            # (B, L, 3, mel)
            out = np.einsum("blkc,ock->blo", frames, w) + sd[n + ".bias"]
        else:
            out = x @ w.T + sd[n + ".bias"]
        out = O.fake_quantize(out.astype(np.float32), np.float32(scale), np.float32(zp), symmetric=False)
        d = np.abs(out - want)
        assert d.max() <= 1.001 * scale + 1e-6, (n, d.max(), scale)
        assert (d < 1e-3 * scale).mean() > 0.995, n


# ------------------------------------------------------------------ host-side algebra of the fused epilogues -----
def test_layernorm_fold_identity():
    """The algebra csrc/engine.cu fold_ln and the EPI_LNA epilogue of csrc/gemm_tc.cu rely on (ssm.py:423-425,
    model.py:229-239, attention.py:306-307):  Linear(LayerNorm(x)) = rstd ((x - c) W'^T) - (mean - c) rstd s + b'
    with W' = W diag(gamma), b' = b + W beta, s[n] = sum_k W'[n, k], for any shift c (the kernel takes the mean of
    the row's first 32 channels).  In fp64 it is exact; in fp32 the shifted form keeps its digits under a common
    offset of the rows where E[x^2] - mean^2 of the raw row does not."""
    rs = np.random.RandomState(11)
    K, N = 192, 40
    W, b = rs.standard_normal((N, K)) / K ** 0.5, rs.standard_normal(N)
    gamma, beta = 1 + 0.3 * rs.standard_normal(K), 0.2 * rs.standard_normal(K)
    Wf, bf, s = W * gamma, b + W @ beta, (W * gamma).sum(1)
    for offset in (0.0, 30.0, 300.0):
        x = rs.standard_normal((7, K)) + offset
        want = O.linear(O.layer_norm(x, gamma, beta), W, b)
        c = x[:, :32].mean(1, keepdims=True)
        xs = x - c
        dm = xs.mean(1, keepdims=True)
        rstd = 1.0 / np.sqrt((xs * xs).mean(1, keepdims=True) - dm * dm + 1e-5)
        got = rstd * (xs @ Wf.T) - dm * rstd * s + bf
        assert np.abs(got - want).max() < 1e-9 * (1 + offset)
        # fp32 statistics: raw sums lose (mean / std)^2 digits, shifted sums do not
        x32 = x.astype(np.float32)
        m_raw = x32.mean(1, dtype=np.float32)
        var_raw = (x32 * x32).mean(1, dtype=np.float32) - m_raw * m_raw
        xs32 = x32 - x32[:, :32].mean(1, keepdims=True, dtype=np.float32)
        dm32 = xs32.mean(1, dtype=np.float32)
        var_sh = (xs32 * xs32).mean(1, dtype=np.float32) - dm32 * dm32
        var64 = x32.astype(np.float64).var(1)                     # of the values the kernel sees
        assert np.abs(var_sh / var64 - 1).max() < 2e-6
        if offset >= 300.0:
            assert np.abs(var_raw / var64 - 1).max() > 1e-4


def test_gate_rows_permutation():
    """csrc/engine.cu gate_perm_row: rows of the stacked gate | local | global projection (attention.py:191-220)
    re-ordered so that each 96-column half of a 192-column tile holds the three 32-row chunks of the SAME 32
    channels — what one epilogue warp of the EPI_GATE variant needs to mix them in registers."""
    C = 192

    def gate_perm_row(rp):
        t, a, j, i = rp // 192, (rp % 192) // 96, (rp % 96) // 32, rp % 32
        return j * C + 64 * t + 32 * a + i

    src = [gate_perm_row(rp) for rp in range(3 * C)]
    assert sorted(src) == list(range(3 * C))
    for rp in range(0, 3 * C, 96):
        chans = [[r % C for r in src[rp + 32 * j: rp + 32 * j + 32]] for j in range(3)]
        assert chans[0] == chans[1] == chans[2] == list(range(chans[0][0], chans[0][0] + 32))
        assert [src[rp + 32 * j] // C for j in range(3)] == [0, 1, 2]
        # output columns of the half: 64 * tile + 32 * half
        assert chans[0][0] == 64 * (rp // 192) + 32 * ((rp % 192) // 96)
