#!/usr/bin/env python
"""Turns an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list of
`bench.py --configs <c> --steps 2 --warmup 1` into (a) a per-kernel table (launches, total ms, DRAM bytes per launch) and
(b) profiles/traffic_<tag>.json, which bench.py reads for roofline.traffic (only for exactly that workload).

    python tools/make_traffic.py gpurun_out/launches_c2.csv phred  k_phred_sum+k_phred_win  <bases_per_gpu>
    python tools/make_traffic.py gpurun_out/launches_c4.csv c3,c4  k_probe_paint            <bases_per_gpu>
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else (v * 1e3 if unit == "s" else v))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)


def main():
    path, tags, kernels, bases = sys.argv[1], sys.argv[2].split(","), sys.argv[3].split("+"), int(sys.argv[4])
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    per = collections.OrderedDict()          # launch id -> dict
    for r in rows:
        if r is hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if not d.get("ID", "").isdigit():
            continue
        e = per.setdefault(d["ID"], {"kernel": d["Kernel Name"].split("(")[0]})
        name = d["Metric Name"]
        if name.startswith("gpu__time"):
            e["ms"] = to_ms(d["Metric Value"], d["Metric Unit"])
        elif name.startswith("dram__bytes"):
            e["bytes"] = e.get("bytes", 0.0) + to_bytes(d["Metric Value"], d["Metric Unit"])
    agg = collections.OrderedDict()
    for e in per.values():
        a = agg.setdefault(e["kernel"], [0, 0.0, 0.0])
        a[0] += 1; a[1] += e.get("ms", 0.0); a[2] += e.get("bytes", 0.0)
    table = [{"kernel": k, "launches": n, "total_ms": round(ms, 3), "dram_bytes_per_launch": int(b / n)} for k, (n, ms, b) in
             sorted(agg.items(), key=lambda x: -x[1][1])]
    sel = [t for t in table if any(k in t["kernel"] for k in kernels)]
    # bytes of one scoring pass = sum over the named kernels of their per-launch bytes (each runs once per step)
    total = sum(t["dram_bytes_per_launch"] for t in sel)
    for tag in tags:
        out = {"scale": 1.0, "bases_per_gpu": bases, "dram_bytes_per_launch": total,
               "source": "%s: dram__bytes_read.sum + dram__bytes_write.sum of %s (ncu launch list of `bench.py --configs ... --steps 2 --warmup 1`, "
                         "per launch)" % (os.path.basename(path), " + ".join(kernels)), "kernels": sel}
        json.dump(out, open(os.path.join(ROOT, "profiles", "traffic_%s.json" % tag), "w"), indent=1)
    json.dump(table[:24], sys.stdout, indent=1)


if __name__ == "__main__":
    main()
