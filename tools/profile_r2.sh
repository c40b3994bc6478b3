#!/bin/bash
# Round-2 profile captures on the GPU box (one GPU).  Outputs under gpurun_out/r2p_*.
P="python tools/traffic_probe.py"
python -c "import bench; print(bench.kernel_src_sha())" > gpurun_out/r2p_kernel_src_sha.txt
$P > gpurun_out/r2p_plain.log 2>&1 || { echo "probe failed"; tail -5 gpurun_out/r2p_plain.log; exit 1; }
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    --csv --log-file gpurun_out/r2p_traffic.csv $P > gpurun_out/r2p_traffic.log 2>&1; echo "traffic rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_sharded.csv \
    $P 65536 8 > gpurun_out/r2p_launches.log 2>&1; echo "launches rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_select_f|k_backprop_f' -s 40 -c 2 \
    -o gpurun_out/r2p_full_c4 -f $P > gpurun_out/r2p_full.log 2>&1; echo "full rc=$?"
# Othello, BASELINE config 4 (4096 trees, n=400, K=4, score utility)
O="python tools/exp_othello.py 4096"
$O > gpurun_out/r2p_oth_plain.log 2>&1 && cat gpurun_out/r2p_oth_plain.log | tail -1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -s 1400 -c 600 --csv \
    --log-file gpurun_out/r2p_oth_launches.csv $O > gpurun_out/r2p_oth_launches.log 2>&1; echo "oth launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_select_ws -s 150 -c 1 -o gpurun_out/r2p_full_oth -f $O > gpurun_out/r2p_oth_full.log 2>&1; echo "oth full select rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_backprop -s 150 -c 1 -o gpurun_out/r2p_full_oth_bp -f $O >> gpurun_out/r2p_oth_full.log 2>&1; echo "oth full backprop rc=$?"
ls -la gpurun_out/ | grep r2p
