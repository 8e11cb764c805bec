"""Turn gpurun_out/*.ncu-rep and launches.csv into small text summaries under profiles/<round>/."""
import collections, csv, glob, os, re, subprocess, sys

out = sys.argv[1] if len(sys.argv) > 1 else "profiles/r1"
os.makedirs(out, exist_ok=True)
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum.per_cycle_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum"]
for rep in sorted(glob.glob("gpurun_out/*.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    name = os.path.basename(rep)[:-8]
    with open(os.path.join(out, name + ".txt"), "w") as fh:
        for r in rows[2:]:
            fh.write("kernel: %s\n" % r[hdr.index("Kernel Name")][:160])
            for w in WANT:
                if w in hdr:
                    fh.write("  %-72s %s %s\n" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    print("wrote", name)
if os.path.exists("gpurun_out/launches.csv"):
    rows = list(csv.reader(open("gpurun_out/launches.csv")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[start:]:
        if len(r) <= mi or r[mi] in ("", "nan"):
            continue
        nm = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")[:80]
        v = float(r[mi].replace(",", "")) / {"ns": 1e6, "nsecond": 1e6, "us": 1e3, "usecond": 1e3, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e6)
        a = agg.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(os.path.join(out, "launch_list_summary.txt"), "w") as fh:
        fh.write("ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 1 --warmup 1 (3 steps incl. e2e leg)\n")
        fh.write("total kernel time %.2f ms in %d launches\n" % (tot, sum(a[0] for a in agg.values())))
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("%-82s n=%4d %9.3f ms %5.1f%%\n" % (k, c, v, 100 * v / tot))
    print("wrote launch_list_summary")
