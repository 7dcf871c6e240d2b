"""Summarise ncu output brought back in gpurun_out/ (read here, no GPU needed).
  python tools/ncu_summary.py launches <launches.csv> [launches_per_step]   -> per-kernel share of the LAST step
  python tools/ncu_summary.py full <report.ncu-rep>                         -> key metrics of each captured launch
"""
import csv, subprocess, sys, re, collections, io


def launches(path, per_step=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r["Metric Name"] == "gpu__time_duration.sum":
            name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("unnamed>::", "").replace("wb::<", "")
            rows.append((name, float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"], r["Block Size"]))
    if per_step:
        rows = rows[-int(per_step):]
    tot = sum(t for _, t, _, _ in rows)
    agg = collections.OrderedDict()
    for n, t, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += t
    print(f"{len(rows)} launches, {tot/1e3:.2f} ms serialised")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:60]:60s} n={c:4d} total={t:10.1f} us avg={t/c:8.1f} us share={100*t/tot:5.1f}%")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_xu", "smsp__inst_executed_pipe_xu", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "lts__t_bytes.sum", "sm__pipe_fma_cycles_active",
        "sm__pipe_alu_cycles_active", "smsp__issue_active.avg.pct", "sm__cycles_active.avg", "smsp__average_warp", "smsp__warp_issue_stalled", "launch__grid_size", "sm__pipe_xu"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        print("==", row[hdr.index("Kernel Name")][:80], "grid", row[hdr.index("Grid Size")], "block", row[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, row):
            if any(k in h for k in KEYS):
                print(f"   {h:90s} {v:>18s} {u}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2])
