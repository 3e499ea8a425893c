"""Per-kernel counts of the SASS mnemonics that prove what the hot kernels run on (cuobjdump -sass of libvasr.so):
UTCHMMA (tcgen05.mma), UTMALDG (TMA load), LDTM / STTM (TMEM load / store), UTCBAR, FFMA2 / FMUL2 / FADD2 (packed
fp32), MUFU, LDGSTS (cp.async), SHFL.   python tools/sass_summary.py > profiles/sass_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "velocity-asr_b200", "velocity_asr", "libvasr.so")
txt = subprocess.run(["cuobjdump", "-sass", so], check=True, capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "LDGSTS", "LDS", "SHFL"]
per, name = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        d = d.replace("(anonymous namespace)::", "").replace("void ", "")
        d = re.sub(r"\(.*", "", d)                               # argument list
        head = re.sub(r"<.*", "", d).split("::")[-1]             # kernel name without namespaces
        targs = d[d.index("<"):] if "<" in d else ""
        name = head + targs
        per[name] = collections.Counter({"_total": 0})
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        per[name][m.group(1)] += 1
        per[name]["_total"] += 1
agg = collections.OrderedDict()
for k, c in per.items():
    base = re.sub(r"<.*", "", k)
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
print("# SASS instruction summary of libvasr.so (sm_100a), per kernel family\n")
print("`python tools/sass_summary.py` on the in-tree build; counts are static instructions summed over a family's template")
print("instantiations.  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR =")
print("tcgen05.commit, FFMA2 / FMUL2 / FADD2 = packed fp32x2, LDGSTS = cp.async.\n")
print("| kernel family | variants | instructions | " + " | ".join(cols) + " |")
print("|---|---:|---:|" + "---:|" * len(cols))
tot = collections.Counter()
for k, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["_total"]):
    print(f"| `{k}` | {n} | {c['_total']} | " + " | ".join(str(c[x]) for x in cols) + " |")
    tot.update(c)
print(f"| **all** | {sum(n for n, _ in agg.values())} | {tot['_total']} | " + " | ".join(str(tot[x]) for x in cols) + " |")
