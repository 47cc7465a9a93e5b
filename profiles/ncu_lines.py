"""Per CUDA source line totals from `ncu --page source --csv --print-source cuda,sass`:
python profiles/ncu_lines.py dump.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
tot_ex = tot_s = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iex = r.index("Instructions Executed"); isamp = r.index("# Samples"); continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "":      # a CUDA source line row (aggregated over its SASS)
        try:
            ex, s = int(r[iex] or 0), int(r[isamp] or 0)
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        agg[key][0] += ex; agg[key][1] += s; agg[key][2] = r[1].strip()[:110]
        tot_ex += ex; tot_s += s
print("total executed warp-instructions %d, samples %d" % (tot_ex, tot_s))
for (f, ln), (ex, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% ex %5.1f%% smp  %-14s:%-4d %s" % (100.0 * ex / max(tot_ex, 1), 100.0 * s / max(tot_s, 1), f, ln, src))
