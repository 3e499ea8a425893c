"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: per-kernel
share of ONE timed step (the launches between two front-end kernels).  Usage:
    python tools/launch_summary.py gpurun_out/launches.csv [step_index] > profiles/rNN_launches.md"""
import collections, csv, re, sys

def main(path, step=3):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "mel_fft_kernel" in n or "reflect_pad" in n]
    i0 = starts[step]
    i1 = starts[step + 1] if step + 1 < len(starts) else len(rows)
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[i0:i1]:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("vasr::", "")
        n = n.replace("<unnamed>::", "").replace("unnamed>::", "")
        key = (n, r["Grid Size"], r["Block Size"])
        t = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
    print(f"step {step}: {i1 - i0} launches, sum of gpu__time_duration = {tot / 1e6:.3f} ms "
          f"(serialised, cold-cache: compare shares, not absolutes)\n")
    print("| us total | share | launches | avg us | kernel | grid | block |\n|---:|---:|---:|---:|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | {v[0]} | {v[1] / v[0] / 1e3:.1f} | `{k[0]}` | {k[1]} | {k[2]} |")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
