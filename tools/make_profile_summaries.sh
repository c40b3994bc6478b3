#!/bin/bash
# gpurun_out/r2p_* (tools/profile_r2.sh) -> the committed summaries under profiles/.  Run in the build container (needs ncu to read the reports).
set -e
python tools/make_traffic_json.py gpurun_out/r2p_traffic.csv "ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none python tools/traffic_probe.py (one search step, whole batch per launch)" | tail -3
python tools/ncu_summary.py gpurun_out/r2p_launches_sharded.csv > profiles/r2_launches_summary_sharded_step_n65536.csv
python tools/ncu_summary.py gpurun_out/r2p_oth_launches.csv > profiles/r2_launches_summary_othello_cfg4.csv
ncu -i gpurun_out/r2p_full_c4.ncu-rep --page raw --csv > /tmp/raw_c4.csv 2>/dev/null; python tools/ncu_full_summary.py /tmp/raw_c4.csv > profiles/r2_ncu_full_c4.csv
ncu -i gpurun_out/r2p_full_oth.ncu-rep --page raw --csv > /tmp/raw_o1.csv 2>/dev/null; ncu -i gpurun_out/r2p_full_oth_bp.ncu-rep --page raw --csv > /tmp/raw_o2.csv 2>/dev/null
python tools/ncu_full_summary.py /tmp/raw_o1.csv > /tmp/o1.csv; python tools/ncu_full_summary.py /tmp/raw_o2.csv > /tmp/o2.csv
python - <<'PY'
import csv
a = list(csv.reader(open('/tmp/o1.csv'))); b = {r[0]: r for r in csv.reader(open('/tmp/o2.csv'))}
w = csv.writer(open('profiles/r2_ncu_full_oth.csv', 'w', newline=''))
for r in a:
    w.writerow(r + (b[r[0]][2:] if r[0] in b else ['']))
PY
