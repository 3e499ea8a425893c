"""`ncu --set full` of the projection launches of ONE config-2 step, summarised per kernel variant and grid.

    ncu --set full --clock-control none -k regex:gemm_tc -s 124 -c 62 -f -o gpurun_out/gemm_full \
        python tools/ncu_gemm_summary.py --run
    python tools/ncu_gemm_summary.py gpurun_out/gemm_full.ncu-rep > profiles/rNN_ncu_gemm_summary.md

--run: three fused transcribe calls at config 2 (62 projection launches each; the capture skips the first two)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "us", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smem/LSU data pipe %", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1),
        ("dram__bytes_read.sum", "DRAM read MB", None),
        ("dram__bytes_write.sum", "DRAM write MB", None),
        ("launch__registers_per_thread", "regs", 1)]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def run():
    sys.path.insert(0, os.path.join(ROOT, "velocity-asr_b200"))
    import torch
    import velocity_asr as va
    torch.manual_seed(0)
    m = va.VELOCITYASR(va.VelocityASRConfig(scan_mode="sequential")).cuda().eval()
    audio = (torch.randn(64, 240000) * 0.1).cuda()
    for _ in range(3):
        m.transcribe(audio)
    torch.cuda.synchronize()


def summarise(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(head)}
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        key = (name, r[ix["Grid Size"]])
        a = agg.setdefault(key, {"n": 0, "sum": collections.defaultdict(float)})
        a["n"] += 1
        for k, _, sc in COLS:
            if k not in ix:
                continue
            try:
                v = float(r[ix[k]].replace(",", ""))
            except ValueError:
                continue
            f = SCALE.get(units[ix[k]], 1.0) if sc is None else TIME.get(units[ix[k]], 1.0) if sc == "time" else sc
            a["sum"][k] += v * f
    print(f"`ncu --set full --clock-control none` of the {len(data)} projection launches of one config-2 step "
          f"(`{os.path.basename(rep)}`); averages per kernel variant and grid.  Template arguments of gemm_tc2_kernel: "
          "<activation, pos-enc, residual, quantised, W-resident, tile columns, epilogue bits (1 folded LayerNorm, "
          "2 argmax partials, 4 gated fusion)>.\n")
    print("| kernel | grid | launches | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|---:|" + "---:|" * len(COLS))
    for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1]["sum"]["gpu__time_duration.sum"]):
        cells = [f"{a['sum'][k] / a['n']:.1f}" for k, _, _ in COLS]
        print(f"| `{name}` | {grid} | {a['n']} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "--run":
        run()
    else:
        summarise(sys.argv[1])
