"""Turn one `ncu --set full` capture of the scan kernel into the tracked record bench.py reads for
`roofline.traffic` (profiles/scan_ncu.json) plus the counters DESIGN.md quotes.

    ncu --set full --clock-control none --import-source on -k regex:scan_ -c 1 -o gpurun_out/scan_full \
        python tools/scan_bench.py --quick
    python tools/ncu_scan_json.py gpurun_out/scan_full.ncu-rep 64 751 384 64 > gpurun_out/scan_ncu.json

(the second command needs the `ncu` binary, so tools/measure_pass.sh runs it on the GPU box; copy the result to
profiles/scan_ncu.json)."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "sm__cycles_active.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(rep, shape):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    rec = {h: (v, u) for h, u, v in zip(head, units, vals)}
    out = {"kernel": rec["Kernel Name"][0].replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", ""),
           "shape": shape, "shape_is": "[B, L, Di, N] of one launch", "source": rep, "metrics": {}}
    for k in KEEP:
        if k in rec:
            v, u = rec[k]
            try:
                out["metrics"][k] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                pass
    for k, name in (("dram__bytes_read.sum", "dram_bytes_read"), ("dram__bytes_write.sum", "dram_bytes_write")):
        v, u = rec[k]
        out[name] = int(round(float(v.replace(",", "")) * SCALE.get(u, 1)))
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], [int(x) for x in sys.argv[2:6]])
