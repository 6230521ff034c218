"""Summarise the ncu launch list of bench.py (gpu__time_duration.sum per launch) for ONE step: kernel families, launch
counts, summed time and share. usage: python scripts/summarize_launches.py gpurun_out/launches_TAG.csv"""
import collections, csv, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ni, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
launches = [(r[ni], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:] if len(r) > vi]
marks = [i for i, (n, _) in enumerate(launches) if "embed_fwd_kernel" in n]
seg = launches[marks[-2]:marks[-1]]


def family(n):
    n = n.replace("void ", "").replace("vb::", "").replace("<unnamed>::", "")
    m = re.match(r"(gemm_tcgen05(?:_2cta)?_kernel)<(\d), (\d), (\d+)(?:, [^>]*)?>", n)
    if m:
        a, b = m.group(2), m.group(3)
        kind = {"00": "fwd (K-major A, K-major B)", "01": "dgrad (B MN-major)", "11": "wgrad (A,B MN-major, fp32 red.add)"}.get(a + b, a + b)
        return f"{m.group(1)} {kind}"
    n = n.split("(")[0]
    if n.startswith("at::") or n.startswith("at_cuda") or "elementwise" in n or "cunn" in n or "cutlass" in n or "nvjet" in n or "cublas" in n:
        return "torch/cuBLAS (heads, index ops, fills)"
    return n.split("<")[0]


agg = collections.OrderedDict()
for n, t in seg:
    f = family(n)
    c = agg.setdefault(f, [0, 0.0])
    c[0] += 1; c[1] += t
tot = sum(v[1] for v in agg.values())
print(f"one bench step (cfg2, B=256, train): {len(seg)} launches, {tot/1e3:.2f} ms summed kernel time under ncu\n")
print("| kernel family | launches | ms | share |\n|---|---|---|---|")
for f, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{f}` | {c} | {t/1e3:.3f} | {100*t/tot:.1f}% |")
