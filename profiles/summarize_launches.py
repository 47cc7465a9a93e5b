"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares as a markdown table.
python profiles/summarize_launches.py gpurun_out/launches_r2_bench.csv [skip_first_N_launches] > profiles/launches_r2.md"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
launches = []
for r in rows:
    if len(r) > 8 and r[0] == "ID":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
        launches.append((int(d["ID"]), d["Kernel Name"], v))
launches = launches[skip:]


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n[:70]


agg = collections.OrderedDict()
for _, n, ms in launches:
    k = short(n)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms (under ncu) | share |")
print("|---|---|---|---|")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f %% |" % (k, c, ms, 100.0 * ms / tot))
print("\n%d launches, %.2f ms in total" % (len(launches), tot))
