"""Summarise an `ncu --set full` report of one layer fwd+bwd (scripts/profile_ncu.sh) into a markdown table
and a traffic json that bench.py reads for roofline.traffic.
usage: python scripts/summarize_ncu.py gpurun_out/prof_layer_TAG.ncu-rep profiles/TAG"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(prefix):
    for i, h in enumerate(hdr):
        if h == prefix:
            return i
    raise KeyError(prefix)


C = {k: col(v) for k, v in dict(
    name="Kernel Name", t="gpu__time_duration.sum", rd="dram__bytes_read.sum", wr="dram__bytes_write.sum",
    tens="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    regs="launch__registers_per_thread", grid="launch__grid_size", block="launch__block_size",
    dram="dram__bytes_read.sum.pct_of_peak_sustained_elapsed", dramw="dram__bytes_write.sum.pct_of_peak_sustained_elapsed", ipc="sm__inst_executed.avg.per_cycle_active",
    warps="sm__warps_active.avg.pct_of_peak_sustained_active").items()}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m[unit]


def to_us(v, unit):
    m = {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}
    return float(v.replace(",", "")) * m[unit]


# order of the library's launches in one layer fwd+bwd (vb_api.cu layer_fwd / layer_bwd)
LABELS = ["fwd qkv GEMM", "fwd attention dropout mask (+transpose)", "fwd attention (tcgen05)", "fwd attn-out GEMM (+bias+dropout+residual)", "fwd LN1", "fwd FFN-up GEMM (+bias+GELU, 2 stores)",
          "fwd FFN-down GEMM (+bias+dropout+residual)", "fwd LN2", "bwd LN2", "bwd FFN-down wgrad", "bwd FFN-down dgrad (*gelu')",
          "bwd colsum (b_inter)", "bwd FFN-up wgrad", "bwd FFN-up dgrad (+addend)", "bwd LN1", "bwd attn-out wgrad", "bwd attn-out dgrad",
          "bwd attention delta", "bwd attention dQ/dK/dV (tcgen05)", "bwd colsum (b_qkv)", "bwd qkv wgrad", "bwd qkv dgrad (+addend)"]
lines = ["| # | launch | kernel | µs | DRAM rd MB | DRAM wr MB | DRAM % | tensor % | IPC | regs | grid×block |", "|---|---|---|---|---|---|---|---|---|---|---|"]
gemm_traffic, gemm_t, tot = [], [], 0.0
for i, r in enumerate(data):
    nm = r[C["name"]].split("(")[0].replace("void ", "")
    t = to_us(r[C["t"]], units[C["t"]])
    rd = to_bytes(r[C["rd"]], units[C["rd"]]); wr = to_bytes(r[C["wr"]], units[C["wr"]])
    tot += t
    if "gemm" in nm:
        gemm_traffic.append(rd + wr); gemm_t.append(t)
    lab = LABELS[i] if len(data) == len(LABELS) else ""
    lines.append(f"| {i} | {lab} | `{nm}` | {t:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {float(r[C['dram']]) + float(r[C['dramw']]):.0f} | {float(r[C['tens']]):.0f} | "
                 f"{float(r[C['ipc']]):.2f} | {r[C['regs']]} | {r[C['grid']]}×{r[C['block']]} |")
lines.append(f"\nsum of the {len(data)} launches: {tot:.0f} µs (ncu times are serialised and cold-cache; compare shares, not absolutes)")
open(out + "_layer_kernels_table.md", "w").write("\n".join(lines) + "\n")
json.dump({"source": rep.split("/")[-1], "workload": "cfg2 B=256 one layer fwd+bwd", "gemm_launches": len(gemm_traffic),
           "gemm_dram_bytes_per_launch_avg": sum(gemm_traffic) / max(1, len(gemm_traffic)),
           "gemm_us_per_launch_avg_under_ncu": sum(gemm_t) / max(1, len(gemm_t))}, open(out + "_traffic.json", "w"), indent=1)
print("\n".join(lines))
