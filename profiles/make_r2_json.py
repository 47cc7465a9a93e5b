"""Turns the ncu CSVs of profiles/r2_final.sh into the small JSON files bench.py reads:
  python profiles/make_r2_json.py traffic  gpurun_out/me_traffic_2160.csv  profiles/me_r2_traffic.json
  python profiles/make_r2_json.py prims    gpurun_out/prims_ncu.csv gpurun_out/prims_single.json  profiles/primitives_r2_ncu.json"""
import csv, json, sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    out = {}
    for r in rows:
        if len(r) > 10 and r[0] == "ID":
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            e = out.setdefault(int(d["ID"]), {"name": d["Kernel Name"]})
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d.get("Metric Unit", "")
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
            e[d["Metric Name"]] = v * mult
    return [out[k] for k in sorted(out)]


if sys.argv[1] == "traffic":
    ls = launches(sys.argv[2])
    phases = {"prechecks": 0.0, "integer_search": 0.0, "subpel": 0.0}
    ms = dict(phases)
    for l in ls:
        n = l["name"]
        b = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        t = l.get("gpu__time_duration.sum", 0.0) / 1e6
        if "k_me_window" in n or ", 2, " in n:
            ph = "integer_search"
        elif ", 1, " in n:
            ph = "prechecks"
        elif ", 3, " in n:
            ph = "subpel"
        else:
            continue
        phases[ph] += b; ms[ph] += t
    out = {"how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over profiles/run_small.py 3840 2160 1 1 1 (one ME stage, c3 workload)",
           "phases": {"c3_3840x2160_8bit_%s" % k: {"dram_bytes_per_step": v, "ncu_ms": ms[k]} for k, v in phases.items()}}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out, indent=1))
else:
    ls = [l for l in launches(sys.argv[2])]
    rows = json.load(open(sys.argv[3]))["rows"]
    base = rows[0]["launch0"]
    out = []
    for r in rows:
        sel = ls[r["launch0"] - base:r["launch1"] - base]
        b = sum(l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in sel)
        out.append({"kernel": r["kernel"], "depth": r["depth"], "dram_bytes": b, "algorithmic_bytes": r["algorithmic_MB"] * 1e6,
                    "launches": [l["name"][:60] for l in sel]})
    json.dump({"how": "ncu dram__bytes_read.sum + dram__bytes_write.sum per primitive launch (profiles/primitive_bench.py --single, 6 stacked 2160p frames)", "rows": out},
              open(sys.argv[4], "w"), indent=1)
    for o in out[:8]:
        print(o)
