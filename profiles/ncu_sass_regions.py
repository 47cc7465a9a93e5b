"""Opcode mix and hot 1 KB code regions of one kernel from `ncu -i rep --page source --csv --print-source sass`:
python profiles/ncu_sass_regions.py dump.csv [region_lo region_hi]   (hex offsets: also list that region's SASS)"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))[2:]
ins = []
for r in rows:
    try:
        ins.append((int(r[0], 16), r[1].strip(), int(r[4] or 0), int(r[5] or 0), int(r[6] or 0)))
    except (ValueError, IndexError):
        pass
tot = sum(i[3] for i in ins); ts = sum(i[2] for i in ins)
print("sass lines %d, executed warp-instructions %d, samples %d, avg active lanes %.1f" % (len(ins), tot, ts, sum(i[4] for i in ins) / max(tot, 1)))
op = collections.Counter(); ops = collections.Counter()
for a, s, smp, ex, th in ins:
    m = re.sub(r'^@!?U?P\d+\s+', '', s).split()[0].split('.')[0]
    op[m] += ex; ops[m] += smp
for k, v in op.most_common(24):
    print("  %-10s %5.1f%% executed %5.1f%% samples" % (k, 100 * v / tot, 100 * ops[k] / ts))
base = ins[0][0]
seg = collections.Counter(); segs = collections.Counter()
for a, s, smp, ex, th in ins:
    seg[(a - base) // 0x400] += ex; segs[(a - base) // 0x400] += smp
print("hot 1 KB code regions (64 instructions each):")
for k, v in sorted(seg.items(), key=lambda kv: -kv[1])[:20]:
    print("  +0x%05x %5.1f%% executed %5.1f%% samples" % (k * 0x400, 100 * v / tot, 100 * segs[k] / ts))
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    for a, s, smp, ex, th in ins:
        if lo <= a - base < hi:
            print("+0x%05x %10d %6d  %s" % (a - base, ex, smp, s))
