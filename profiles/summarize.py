"""Summarise ncu artefacts brought back in gpurun_out/ into text files under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
    python profiles/summarize.py full gpurun_out/prof_ss_simt.ncu-rep > profiles/r1_ss_simt_full.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"][:90]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, row["Metric Unit"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel totals from %s (ncu gpu__time_duration.sum, cold-cache, serialised: compare SHARES)" % path)
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-92s n=%4d total=%12.1f %s avg=%10.1f share=%5.1f%%" % (k, a[0], a[1], a[2], a[1] / a[0], 100 * a[1] / tot))


def full(path):
    """path: a .ncu-rep, or the `ncu -i … --page raw --csv` text already made on the GPU box (capture.sh)."""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    print("# selected metrics from %s (ncu --set full --clock-control none)" % path)

    def num(row, name):
        return float(row[hdr.index(name)].replace(",", "")) if name in hdr and row[hdr.index(name)] else 0.0

    def scale(name, to):   # ncu picks the unit per column: normalise to bytes / ns
        u = units[hdr.index(name)] if name in hdr else ""
        return {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(u, 1.0)

    for row in r[2:]:
        if "Kernel Name" in hdr and ("at::" in row[hdr.index("Kernel Name")] or "cub::" in row[hdr.index("Kernel Name")]):
            continue    # torch's own kernels of the probe script (table normalisation, fills)
        for i in idx:
            print("  %s = %s %s" % (hdr[i], row[i], units[i]))
        rd = num(row, "dram__bytes_read.sum") * scale("dram__bytes_read.sum", "byte")
        wr = num(row, "dram__bytes_write.sum") * scale("dram__bytes_write.sum", "byte")
        ns = num(row, "gpu__time_duration.sum") * scale("gpu__time_duration.sum", "ns")
        l2 = num(row, "lts__t_bytes.sum") * scale("lts__t_bytes.sum", "byte")
        if ns:
            print("  derived: DRAM traffic %.3f MB -> %.1f GB/s achieved%s" % (
                (rd + wr) / 1e6, (rd + wr) / ns, "; L2 traffic %.3f MB -> %.1f GB/s" % (l2 / 1e6, l2 / ns) if l2 else ""))
        print("  --")




def source(path, top=45):
    """Top SASS lines by stall samples and by executed instructions (ncu --page source)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == hdr[0]:
            if body:
                break          # only the first captured launch
            continue
        body.append(r)
    tot_s = sum(int(r[ci["# Samples"]]) for r in body) or 1
    tot_i = sum(int(r[ci["Instructions Executed"]]) for r in body) or 1
    print("# %s: %d SASS lines, %d samples, %d warp-instructions executed" % (path, len(body), tot_s, tot_i))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ci[h]]) for r in body) for h in stall_cols}
    print("# stall mix:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / tot_s) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    print("# --- in program order, lines with >=0.4%% of samples or >=0.6%% of instructions")
    for n, r in enumerate(body):
        s, i = int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]])
        if s >= 0.004 * tot_s or i >= 0.006 * tot_i:
            st = sorted(((int(r[ci[h]]), h[6:]) for h in stall_cols), reverse=True)[:2]
            print("%5d %-70s samp=%5.1f%% inst=%5.1f%% thr/inst=%4s  %s" % (
                n, r[ci["Source"]].strip()[:70], 100.0 * s / tot_s, 100.0 * i / tot_i, r[ci["Avg. Threads Executed"]],
                " ".join("%s:%d" % (b, a) for a, b in st if a)))


if __name__ == "__main__":
    {"launches": launches, "full": full, "source": source}[sys.argv[1]](sys.argv[2])
