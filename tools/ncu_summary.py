"""Condense an `ncu --csv --log-file` launch list (any --metrics set) into one row per kernel: launches, mean of every metric.
    python tools/ncu_summary.py gpurun_out/traffic_r1g.csv [first_id last_id] > profiles/<name>.csv"""
import collections, csv, sys

UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}


def main():
    path = sys.argv[1]
    lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 60)
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        i = int(r[ix["ID"]])
        if not lo <= i <= hi:
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        per.setdefault((i, name), {})[r[ix["Metric Name"]]] = v * UNIT.get(u, 1.0)
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        for k, v in m.items():
            agg.setdefault(name, collections.OrderedDict()).setdefault(k, []).append(v)
    metrics = sorted({k for m in agg.values() for k in m})
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches"] + [f"mean {m}" + (" [ns]" if "time" in m else " [bytes]" if "bytes" in m else "") for m in metrics])
    for name, m in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
        n = max(len(v) for v in m.values())
        w.writerow([name, n] + [f"{sum(m[k]) / len(m[k]):.6g}" if k in m else "" for k in metrics])


if __name__ == "__main__":
    main()
