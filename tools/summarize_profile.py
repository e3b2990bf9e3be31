"""Turns the files tools/profile.sh leaves in gpurun_out/ into the tables of profiles/<tag>_summary.md:
launch list of one pass (between two k_pre_s8 launches) aggregated by kernel, and the key metrics of the full captures."""
import collections, csv, re, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"

rows = []
with open(f"{src}/{tag}_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        v = float(x["Metric Value"].replace(",", ""))
        rows.append((x["Kernel Name"], v / 1000 if x["Metric Unit"] in ("ns", "nsecond") else v))
idx = [i for i, (k, _) in enumerate(rows) if "k_pre_s8" in k]
a, b = idx[1], idx[2]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows[a:b]:
    k = re.sub(r"^void ", "", re.sub(r"\(.*", "", k))
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"## Launch list: one pass = {b - a} launches, {tot:.0f} us serialised (ncu)")
print("| kernel | launches | us | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "lts__t_bytes.sum"]
for n in ["resblock", "stream", "stream3d", "costvol"]:
    try:
        rd = list(csv.reader(open(f"{src}/{tag}_{n}_raw.csv")))
    except OSError:
        continue
    hdr = rd[0]
    print(f"\n## full capture: {n}")
    for w in ["Kernel Name"] + want:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} [{rd[1][i]}] | " + " / ".join(r[i] for r in rd[2:]) + " |")
