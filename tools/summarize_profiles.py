"""Turn the ncu artifacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

usage: python tools/summarize_profiles.py <tag> [launch_csv] [step_csv] [full.ncu-rep] [bench.log]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launch_csv = sys.argv[2] if len(sys.argv) > 2 else f"{ROOT}/gpurun_out/launches_{tag}.csv"
step_csv = sys.argv[3] if len(sys.argv) > 3 else f"{ROOT}/gpurun_out/step_launches.csv"
full_rep = sys.argv[4] if len(sys.argv) > 4 else f"{ROOT}/gpurun_out/prof_{tag}.ncu-rep"
bench_log = sys.argv[5] if len(sys.argv) > 5 else f"{ROOT}/gpurun_out/bench.log"
out = f"{ROOT}/profiles"


def us(v, u):
    v = float(v.replace(",", ""))
    return v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v


def rows_of(path):
    return list(csv.DictReader(l for l in open(path) if not l.startswith("==")))


if os.path.exists(launch_csv):
    agg = collections.OrderedDict()
    for r in rows_of(launch_csv):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r["Kernel Name"][:100], [0, 0.0])
        a[0] += 1
        a[1] += us(r["Metric Value"], r["Metric Unit"])
    tot = sum(a[1] for a in agg.values())
    with open(f"{out}/{tag}_launch_summary.txt", "w") as f:
        f.write(f"# {tag} -- ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu --no-graph` "
                "(--metrics gpu__time_duration.sum --clock-control none)\n"
                "# cold-cache, serialised launches: compare SHARES, not absolutes. Covers the device-timed steps, the chunked\n"
                "# e2e steps (smaller launches), the profiled reps, the L2 flushes (at::...) and the coverage passes.\n")
        f.write("%-100s %6s %12s %9s %7s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-100s %6d %12.1f %9.1f %6.1f%%\n" % (k, n, t, t / n, 100 * t / tot))
    print("wrote launch summary,", len(agg), "kernels")

if os.path.exists(step_csv):
    by = collections.OrderedDict()
    for r in rows_of(step_csv):
        by.setdefault((r["ID"], r["Kernel Name"][:60]), {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    tot = sum(us(*m["gpu__time_duration.sum"]) for m in by.values())
    with open(f"{out}/{tag}_step_launches.tsv", "w") as f:
        f.write(f"# {tag} -- ONE clustering pass (20M-signal 30X set, eps 500, m 3) of tools/profile_target.py, kernel by kernel\n"
                "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm/dram %% --clock-control none\n")
        f.write("id\tkernel\tus\tshare\tdram_read_B\tdram_write_B\tsm_pct\tdram_pct\n")
        for (i, k), m in by.items():
            t = us(*m["gpu__time_duration.sum"])
            g = lambda n: m.get(n, ("0", ""))[0].replace(",", "")
            f.write("%s\t%s\t%.1f\t%.1f%%\t%s\t%s\t%.1f\t%.1f\n" % (
                i, k, t, 100 * t / tot, g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
                float(g("sm__throughput.avg.pct_of_peak_sustained_elapsed")),
                float(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))))
        f.write("# total %.1f us\n" % tot)
    print("wrote step launches, total %.1f us" % tot)

raw_csv = f"{ROOT}/gpurun_out/prof_{tag}_raw.csv"     # `ncu -i ... --page raw --csv` run on the box (the report is too big to bring back)
if os.path.exists(full_rep) or os.path.exists(raw_csv):
    if os.path.exists(raw_csv):
        raw = open(raw_csv).read()
    else:
        raw = subprocess.run(["ncu", "-i", full_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
    idx = [hdr.index(w) for w in want if w in hdr]
    traffic = collections.OrderedDict()
    with open(f"{out}/{tag}_full_capture_summary.tsv", "w") as f:
        f.write(f"# {tag} -- ncu --set full --clock-control none, tools/profile_target.py (20M-signal 30X set: 3rd clustering pass "
                "onwards; 61.8M reads coverage x2; 250 Mbp GC; candidate aggregation; 61.8M-bin ploidy medians)\n")
        f.write("\t".join(hdr[i] for i in idx) + "\n")
        for r in rows[2:]:
            f.write("\t".join((r[i][:60] + (" " + units[i] if units[i] else "")) for i in idx) + "\n")
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "Tbyte": 1e12}
            t = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
            traffic.setdefault(name, []).append(t)
    # per-launch DRAM traffic of the named kernels (first launch of each = the x-axis / big one)
    json.dump({k: {"first_launch_bytes": v[0], "launches": len(v), "mean_bytes": sum(v) / len(v)} for k, v in traffic.items()},
              open(f"{out}/{tag}_traffic.json", "w"), indent=1)
    print("wrote full capture summary,", len(rows) - 2, "launches")

if os.path.exists(bench_log):
    for line in open(bench_log):
        if line.startswith("{"):
            json.dump(json.loads(line), open(f"{out}/{tag}_bench.json", "w"), indent=1)
            print("wrote bench json")
