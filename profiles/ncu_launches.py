"""Markdown table of one analysed frame from an `ncu --metrics ... --csv --log-file` launch list:
python profiles/ncu_launches.py launches.csv [first_launch_id]"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(int(r[ii]), {"name": r[ki]})
    v = float(r[vi].replace(',', '')); u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
    if r[mi].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r[mi]] = v
ids = list(L)
first = int(sys.argv[2]) if len(sys.argv) > 2 else None
if first is None:      # last occurrence of k_build_me_jobs starts the last step
    first = max(i for i in ids if "k_build_me_jobs" in L[i]["name"])
step = [(i, L[i]) for i in ids if i >= first]
tot = sum(d["gpu__time_duration.sum"] for i, d in step)
print("| # | kernel | ms | share | DRAM read MB | DRAM write MB | warp-instr (M) | issue active % | l1tex % | warps active % | i-cache hit % |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
g = lambda d, k: d.get(k, 0.0)
for i, d in step:
    nm = re.sub(r'\(.*', '', d["name"]).replace("void ", "").replace("unsigned char", "u8").replace("(int)", "")
    print("| %d | `%s` | %.3f | %.1f%% | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f |" % (
        i, nm, d["gpu__time_duration.sum"], 100 * d["gpu__time_duration.sum"] / tot, g(d, "dram__bytes_read.sum") / 1e6, g(d, "dram__bytes_write.sum") / 1e6,
        g(d, "smsp__inst_executed.sum") / 1e6, g(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g(d, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), g(d, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(d, "sm__icc_request_hit_rate.pct")))
print("\nsum of the step's launches: %.2f ms" % tot)
