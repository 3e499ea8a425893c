"""Small end-to-end pass for compute-sanitizer (memcheck): every kernel of the path at tiny sizes, ragged tails."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import velocity_asr as va
import fixtures_util as FU
torch.manual_seed(0)
for mode in ("sequential", "parallel"):
    m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode=mode)).cuda().eval()
    audio = FU.synth_audio(3, 4000 + 37, seed=5).cuda()
    mel = va.compute_mel_spectrogram(audio)
    logits, feats = m(mel, return_features=True)
    toks = m.transcribe(audio)
    ts = va.ctc_greedy_decode_with_timestamps(logits)
    print(mode, mel.shape, logits.shape, [len(t) for t in toks], torch.isfinite(logits).all().item())
x = torch.randn(2, 37, 384, device="cuda"); dt = torch.rand(2, 37, 384, device="cuda")
A = -torch.arange(1, 65, device="cuda", dtype=torch.float32); Bm = torch.randn(2, 37, 64, device="cuda"); Cm = torch.randn(2, 37, 64, device="cuda")
for mode in ("sequential", "parallel"):
    print(mode, va.selective_scan(x, dt, A, Bm, Cm, torch.ones(384, device="cuda"), z=x, scan_mode=mode).abs().max().item())
    print(mode, "generic A", va.selective_scan(x, dt, A * 1.03, Bm, Cm, None, scan_mode=mode).abs().max().item())
q = va.prepare_model_for_qat(va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential")).cuda().eval())
va.calibrate_model(q, [mel]); print("quant", q(mel).abs().max().item())
torch.cuda.synchronize(); print("done")
