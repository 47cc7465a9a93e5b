"""Print selected metrics from an `ncu --page raw --csv` dump: python profiles/ncu_pick.py file.csv [substr ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1] if len(rows) > 2 else [""] * len(hdr)
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "registers_per_thread", "warps_active.avg.pct",
                         "sm__throughput.avg.pct", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate.pct",
                         "issue_stalled", "shared_mem", "grid_size", "block_size", "local_load", "local_store", "thread_inst_executed_per_inst",
                         "occupancy", "dram__throughput.avg.pct", "issue_active.avg.pct", "inst_executed_pipe"]
for r in rows[2:] if len(rows) > 2 else rows[1:]:
    print("==", r[hdr.index("Kernel Name")][:60])
    for h, u, v in zip(hdr, units, r):
        if any(p in h for p in pats):
            print("  %-95s %-12s %s" % (h[:95], u, v))
