"""ncu metrics CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) of tools/traffic_probe.py ->
profiles/dram_traffic.json, keyed by the sha of the kernel sources it was captured from (bench.py refuses it when stale).
    python tools/make_traffic_json.py gpurun_out/<csv> "<ncu command line>" """
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
ix = {h: i for i, h in enumerate(rows[0])}
per = collections.OrderedDict()
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("az::", "")
    base = name.split("<")[0]
    v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
    per.setdefault((int(r[ix["ID"]]), base), {})[r[ix["Metric Name"]]] = v * UNIT.get(u, 1.0)
agg = {}
for (_, base), m in per.items():
    a = agg.setdefault(base, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "gpu_time_ns": 0.0})
    a["launches"] += 1
    a["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
    a["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
    a["gpu_time_ns"] += m.get("gpu__time_duration.sum", 0.0)
sha_here = bench.kernel_src_sha()
sha_file = os.path.join(os.path.dirname(os.path.abspath(sys.argv[1])), "r2p_kernel_src_sha.txt")
sha_box = open(sha_file).read().strip() if os.path.exists(sha_file) else None
if sha_box and sha_box != sha_here:
    raise SystemExit(f"the capture was taken from kernel sources {sha_box}, the tree now holds {sha_here}: re-capture")
out = {"kernel_src_sha": sha_box or sha_here, "captured_from": sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1]),
       "note": "sums over the launches of ONE search step (51 iterations, whole batch of 65 536 trees per launch); bench.py divides by `launches`",
       "kernels": agg}
json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
for k, a in agg.items():
    print(k, a["launches"], "launches, DRAM read %.1f MB + write %.1f MB per launch, %.1f us per launch (serialised, under ncu)" % (
        a["dram_read_bytes"] / a["launches"] / 1e6, a["dram_write_bytes"] / a["launches"] / 1e6, a["gpu_time_ns"] / a["launches"] / 1e3))
