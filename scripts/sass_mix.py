"""Opcode mix + stall-sample share per opcode for one kernel of an ncu report (source page, SASS view).
usage: python scripts/sass_mix.py report.ncu-rep kernel-regex [top]"""
import collections, csv, io, re, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Instructions Executed" in r)
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows if len(r) == len(hdr) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[isamp] or 0) for r in data)
print(f"{len(data)} SASS instructions, {tot} warp-instructions executed, {tots} samples")
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    o = m.group(2).split(".")[0] if m else "?"
    op[o] += int(r[ie]); ops[o] += int(r[isamp] or 0)
for o, c in op.most_common(top):
    print(f"{o:10s} {c:12d} {100*c/tot:5.1f}%   samples {100*ops[o]/max(1,tots):5.1f}%")
