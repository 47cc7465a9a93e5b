"""Summarise an `ncu --page source --csv` dump (SASS view): opcode histogram by executed instructions and
by stall samples, plus the hottest instructions.  python profiles/ncu_sass_top.py file.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops_ex, ops_samp = collections.Counter(), collections.Counter()
tot_ex = tot_s = 0
items = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ex, s = int(r[iex] or 0), int(r[isamp] or 0)
    ops_ex[op] += ex; ops_samp[op] += s; tot_ex += ex; tot_s += s
    items.append((s, ex, src))
print("total warp-instructions executed:", tot_ex, " samples:", tot_s)
print("-- by executed instructions")
for op, n in ops_ex.most_common(22): print("  %-12s %12d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot_ex, 100.0 * ops_samp[op] / max(tot_s, 1)))
print("-- hottest instructions by stall samples")
for s, ex, src in sorted(items, reverse=True)[:25]: print("  %7d %10d  %s" % (s, ex, src[:90]))
