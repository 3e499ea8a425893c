#!/usr/bin/env python
"""bench.py — RTFx of the VELOCITY-ASR v2 inference path (PCM -> log-mel -> model -> greedy CTC).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference's CPU path

Workload (BASELINE.json configs[1]): 64 x 15 s synthetic 16 kHz utterances per GPU, FP32,
random-init weights (torch.manual_seed(0)), audio randn*0.1 (seed 1234).  Weak scaling: every
rank processes its own 64 utterances (N = 8 is configs[2], 512 utterances); no collective on the
data path.  A "step" is one pass of the whole path over the rank's batch.

  value  = audio-seconds per wall-second with PCM already resident in HBM (vasr_transcribe),
           timed with CUDA events, max over ranks.
  e2e    = same metric through the public API with HOST buffers (VELOCITYASR.transcribe_batches over
           pinned tensors): host->device copy of the PCM and device->host copy of the token ids
           inside the timed region, every step, overlapped with the previous / next step's kernels;
           e2e.single_call is one blocking VELOCITYASR.transcribe(host tensor) per step.
  roofline = the selective-scan kernel (the kernel BASELINE.json's metric names): algorithmic
           bytes per launch (4*(4*Di + 2*N) = 6,656 B per token-layer with
           the silu(z) gate fused) / average launch duration from CUDA events on the launch stream.
  cpu_baseline = the reference's own PyTorch CPU path (baseline/_ref, kind "reference"; falls
           back to the numpy oracle port) on a bounded sample, rank 0 at N = 1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "velocity-asr_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SR = 16000
UTT_SECONDS = 15
BATCH_PER_GPU = 64
SCAN_MODE = "sequential"      # true recurrence; see DESIGN.md (the reference's 'parallel' needs ~1 GB/utt on CPU)
METRIC = "rtfx"
UNIT = "audio-seconds/second"
WORKLOAD = "configs[1]: batch 64 x 15 s synthetic 16 kHz utterances per GPU, FP32, mel->SSM->attention->CTC greedy"


def synth_audio(batch, samples, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, samples, generator=g) * 0.1


# ----------------------------------------------------------------------------- clocks -----
class ClockSampler:
    """nvidia-smi sampled every 20 ms (B200_PROFILING.md); only samples that arrive between mark_begin()
    and mark_end() — the timed region — are reported, so idle clocks before / after it do not count."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.15)               # let the first samples arrive before the region starts
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[0]))
                    mx.append(float(r[1]))
                    for n, v in zip(names, r[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
                except Exception:
                    pass
            return sorted(sm), mx, reasons
        inside = [x for x in self.rows if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30) + 0.02]
        sm, mx, reasons = collect(inside)
        where = "timed region"
        if not sm:                          # region shorter than the sampling period: nearest samples
            sm, mx, reasons = collect(self.rows)
            where = "whole run (timed region shorter than the sampling period)"
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL announces its version on the first
# communicator), so file descriptor 1 points at stderr while the benchmark runs and is restored for the line.
_REAL_STDOUT = None


def quiet_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- reference ---
def load_reference_impl():
    """(kind, module): the unmodified reference if it travelled with the repo, else None."""
    try:
        from refload import load_reference
        ref = load_reference()
        if ref is not None:
            return "reference", ref
    except Exception:
        pass
    return "port", None


def cpu_path_rtfx(n_utt, seconds, repeats, scan_mode):
    """Reference CPU path (compute_mel_spectrogram -> model -> ctc_greedy_decode, eval, no_grad,
    all host threads) on n_utt utterances; falls back to the numpy oracle port."""
    import numpy as np
    import torch
    kind, ref = load_reference_impl()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    audio = synth_audio(n_utt, SR * seconds, 1234)
    best = None
    if ref is not None:
        torch.manual_seed(0)
        model = ref.VELOCITYASR(ref.VelocityASRConfig(scan_mode=scan_mode)).eval()

        def run():
            with torch.no_grad():
                mel = ref.compute_mel_spectrogram(audio)
                return ref.ctc_greedy_decode(model(mel))
    else:
        import velocity_asr as va
        import velocity_oracle as O
        torch.manual_seed(0)
        sd = {k: v.numpy() for k, v in va.VELOCITYASR(va.VelocityASRConfig()).state_dict().items()}
        a = audio.numpy()

        def run():
            return O.transcribe(a, sd, dict(scan_mode=scan_mode), dtype=np.float32)
    run()                                   # warm-up (lazy init, thread pools)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": n_utt * seconds / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{n_utt} x {seconds} s utterances, scan_mode={scan_mode}, best of {repeats} after 1 warm-up, "
                      f"{best:.2f} s per pass", "seconds_per_pass": best, "times": times}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_utt = args.ref_utts
    t_all = []
    res = None
    kind, ref = load_reference_impl()
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    audio = synth_audio(n_utt, SR * UTT_SECONDS, 1234)
    if ref is not None:
        torch.manual_seed(0)
        model = ref.VELOCITYASR(ref.VelocityASRConfig(scan_mode=SCAN_MODE)).eval()

        def step():
            with torch.no_grad():
                return ref.ctc_greedy_decode(model(ref.compute_mel_spectrogram(audio)))
    else:
        import numpy as np
        import velocity_asr as va
        import velocity_oracle as O
        torch.manual_seed(0)
        sd = {k: v.numpy() for k, v in va.VELOCITYASR(va.VelocityASRConfig()).state_dict().items()}
        a = audio.numpy()

        def step():
            return O.transcribe(a, sd, dict(scan_mode=SCAN_MODE), dtype=np.float32)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    total = time.perf_counter() - t0
    value = args.steps * n_utt * UTT_SECONDS / total
    sample = (f"each step = {n_utt} of the workload's {BATCH_PER_GPU} utterances (x {UTT_SECONDS} s), "
              f"scan_mode={SCAN_MODE}, torch CPU, {torch.get_num_threads()} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": n_utt, "seconds_per_utterance": UTT_SECONDS,
                   "scan_mode": SCAN_MODE, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- own arm ----
def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_own_arm(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import velocity_asr as va
    from velocity_asr import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, S = args.batch, SR * UTT_SECONDS
    torch.manual_seed(0)
    model = va.VELOCITYASR(va.VelocityASRConfig(scan_mode=SCAN_MODE)).to(dev).eval()
    eng = model._engine(dev)
    lib = eng.lib
    # four distinct input sets (4 x 61 MB > the 126 MB L2) rotated across steps; each step also
    # streams > 1 GB of activations, so nothing survives in L2 from one step to the next.
    n_sets = 4
    host = [synth_audio(B, S, 1234 + rank * 16 + i).pin_memory() for i in range(n_sets)]
    devb = [h.to(dev) for h in host]
    T = 1 + S // 160
    L = (T + 1) // 2
    tokens = torch.empty(B, L, dtype=torch.int32, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def step_dev(i):
        _native.check(lib.vasr_transcribe(eng.handle, _native.ptr(devb[i % n_sets]), B, S, _native.ptr(tokens),
                                          _native.ptr(lens), sp))

    # ---- device-resident throughput
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.vasr_kernel_launches(eng.handle)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record(stream)
    for i in range(args.steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    sampler.mark_end()
    dev_ms = e0.elapsed_time(e1)
    launches = lib.vasr_kernel_launches(eng.handle) - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- scan kernel timing (separate pass so event pairs do not perturb the number above)
    lib.vasr_set_timing(eng.handle, 1)
    scan_ms, n_scan = [], 0
    for i in range(max(3, min(args.steps, 10))):
        step_dev(i)
        torch.cuda.synchronize()
        sm, sl, tot = ctypes.c_float(), ctypes.c_int32(), ctypes.c_float()
        _native.check(lib.vasr_last_timing(eng.handle, ctypes.byref(sm), ctypes.byref(sl), ctypes.byref(tot)))
        scan_ms.append((sm.value, tot.value))
        n_scan = sl.value
    lib.vasr_set_timing(eng.handle, 0)

    # ---- end to end through the public API, host buffers in, token lists out.
    # (a) the streaming call: VELOCITYASR.transcribe_batches(iterable of pinned host batches) copies batch i+1
    #     host->device and batch i-1's token ids device->host while batch i computes; every step's H2D and D2H
    #     are inside the timed region.  (b) one blocking transcribe(host batch) call per step, for reference.
    for _ in model.transcribe_batches(host[i % n_sets] for i in range(max(2, args.warmup))):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for out in model.transcribe_batches(host[i % n_sets] for i in range(args.steps)):
        n_out += len(out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert n_out == args.steps * B
    barrier()
    for i in range(max(1, args.warmup)):
        model.transcribe(host[i % n_sets])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = model.transcribe(host[i % n_sets])
    torch.cuda.synchronize()
    e2e1_s = time.perf_counter() - t0
    barrier()

    times = torch.tensor([dev_ms, e2e_s * 1e3, e2e1_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e1_ms = float(times[0]), float(times[1]), float(times[2])

    if rank == 0:
        audio_s = world * B * UTT_SECONDS * args.steps
        cfg = model.config
        di = cfg.d_model * cfg.ssm_expand_ratio
        # local scan launches come first in each step; the 2 global launches (93 tokens) are tiny
        local_bytes = B * L * 4 * (4 * di + 2 * cfg.ssm_state_dim)
        K1 = min(max(64, L // 8), L)
        global_bytes = B * K1 * 4 * (4 * di + 2 * cfg.global_ssm_state_dim)
        scan_total_ms = sum(s for s, _ in scan_ms) / len(scan_ms)
        step_total_ms = sum(t for _, t in scan_ms) / len(scan_ms)
        all_bytes = cfg.ssm_layers * local_bytes + cfg.global_ssm_layers * global_bytes
        # per-launch average over the local-layer launches: apportion the measured total by bytes
        local_ms = scan_total_ms * (cfg.ssm_layers * local_bytes / all_bytes) / cfg.ssm_layers
        peak, peak_src = hbm_peak()
        achieved = local_bytes / (local_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": audio_s / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world,
                       "seconds_per_utterance": UTT_SECONDS, "tokens_per_utterance": L, "scan_mode": SCAN_MODE,
                       "parallelism": f"dp{world} (utterance shards, no data-path collective)",
                       "l2": "4 rotating input sets (244 MB) + >1 GB of activations per step: larger than L2"},
            "e2e": {"value": audio_s / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": B * S * 4, "d2h_bytes_per_step": B * L * 4 + B * 4,
                    "api": "VELOCITYASR.transcribe_batches(pinned host batches) -> List[List[int]] per batch "
                           "(H2D of batch i+1 and D2H of batch i-1 overlap the kernels of batch i)",
                    "single_call": {"value": audio_s / (e2e1_ms * 1e-3), "ms_per_step": e2e1_ms / args.steps,
                                    "api": "VELOCITYASR.transcribe(pinned host tensor), one blocking call per step"}},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "scan_seq_kernel<LPR=4, 4 warps, 2 rows/lane, structured A> (8 local SSM layers)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape, from the
                         # ncu --set full capture summarised in profiles/r01_ncu_full_summary.md
                         "traffic": 303103232 if (B, L, di, cfg.ssm_state_dim) == (64, 751, 384, 64) else None,
                         "algorithmic_bytes_per_launch": local_bytes, "avg_launch_ms": local_ms,
                         "scan_launches_per_step": n_scan, "scan_ms_per_step": scan_total_ms,
                         "scan_share_of_step": scan_total_ms / step_total_ms,
                         "note": "fp32-issue bound, not HBM bound (DESIGN.md): 24,576 state updates x 3 fp32 ops "
                                 "per token-layer"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_path_rtfx(args.cpu_utts, UTT_SECONDS, 2, SCAN_MODE)
            except Exception as exc:  # the baseline must never sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {exc!r}"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="utterances per GPU")
    ap.add_argument("--cpu-utts", type=int, default=16, help="utterances in the cpu_baseline sample")
    ap.add_argument("--ref-utts", type=int, default=8, help="utterances per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
