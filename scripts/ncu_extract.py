"""Turn the raw ncu CSVs of scripts/ncu_r02.sh into the summaries committed under profiles/:
   python scripts/ncu_extract.py gpurun_out profiles r02
 <tag>_launches_one_step.csv  every launch of the LAST captured training step: kernel, grid, block, time, DRAM bytes
 <tag>_launch_summary.csv     the same aggregated per kernel with its share of the step
 <tag>_kernels_ncu_full_extract.csv  --set full: per launch tensor-pipe %, L2 %, DRAM bytes, issue %, ALU %, occupancy, registers"""
import collections
import csv
import sys

src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "")
    return n.split("(")[0]


# ---- launch list (one row per metric per launch in ncu's CSV)
rows = [r for r in csv.reader(open(f"{src}/{tag}_launches_raw.csv")) if len(r) > 10]
hdr = rows[0]
I = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value")}
launches = collections.OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(int(r[I["ID"]]), {"kernel": short(r[I["Kernel Name"]]), "grid": r[I["Grid Size"]], "block": r[I["Block Size"]]})
    d[r[I["Metric Name"]]] = float(r[I["Metric Value"]].replace(",", ""))
seq = list(launches.values())
# the last training step starts at the last weight-cast kernel
starts = [i for i, d in enumerate(seq) if d["kernel"].startswith("krsc_to_bf16")]
step = seq[starts[-1] - 1:] if starts else seq   # the statistics fill precedes the cast
tot = sum(d["gpu__time_duration.sum"] for d in step)
with open(f"{dst}/{tag}_launches_one_step.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB"])
    for d in step:
        w.writerow([d["kernel"], d["grid"], d["block"], round(d["gpu__time_duration.sum"] / 1e3, 2), round(d.get("dram__bytes_read.sum", 0) / 1e6, 2),
                    round(d.get("dram__bytes_write.sum", 0) / 1e6, 2)])
agg = collections.OrderedDict()
for d in step:
    a = agg.setdefault(d["kernel"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
with open(f"{dst}/{tag}_launch_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches_per_step", "time_us_sum", "share_pct", "dram_read_MB", "dram_write_MB"])
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, a[0], round(a[1] / 1e3, 1), round(100 * a[1] / tot, 1), round(a[2] / 1e6, 1), round(a[3] / 1e6, 1)])
print(f"one step: {len(step)} launches, {tot / 1e6:.3f} ms serialised")

# ---- --set full extract
rows = list(csv.reader(open(f"{src}/{tag}_full_raw.csv")))
hdr = rows[0]
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("gpu__time_duration.sum", "time_us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"), ("dram__bytes_read.sum", "dram_read_MB"),
        ("dram__bytes_write.sum", "dram_write_MB"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp_instructions")]
with open(f"{dst}/{tag}_kernels_ncu_full_extract.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([c[1] for c in cols])
    for r in rows[2:]:
        out = []
        for k, _n in cols:
            v = r[hdr.index(k)] if k in hdr else ""
            out.append(short(v) if k == "Kernel Name" else v)
        w.writerow(out)
print(f"full extract: {len(rows) - 2} launches")
