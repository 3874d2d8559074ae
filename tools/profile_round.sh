#!/bin/bash
# One gpurun command for a round's single-GPU evidence (about 5 GPU-minutes):
#     gpurun --timeout 900 -- 'bash tools/profile_round.sh r02'
# Writes everything under gpurun_out/<tag>_*; summarise the .ncu-rep files here with tools/ncu_summary.py and copy what
# should be judged into profiles/.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
set -x
timeout 300 python -m pytest tests -m gpu -q > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench.err
timeout 60 python tools/shard_probe.py > $O/${TAG}_shard_probe.txt 2>&1
timeout 60 python tools/gnn_batch_once.py > $O/${TAG}_gnn_batch.txt 2>&1
timeout 60 python tools/k1_vs_iteration.py > $O/${TAG}_k1_vs_iteration.txt 2>&1
# launch list of bench.py's step (kernel shares), then one full capture per kernel family that changed this round
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aco_|tsp_|hadamard|knn_refresh' -s 9 -c 40 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:aco_knn -s 4 -c 1 -o $O/${TAG}_k1_knn \
    python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tsp_update_kernel -s 4 -c 1 -o $O/${TAG}_k2_update \
    python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:tsp_update_row -s 2 -c 1 -o $O/${TAG}_k2_update_rows \
    python tools/update_once.py 16384 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:gnn_forward_kernel -s 1 -c 1 -o $O/${TAG}_k3_gnn_batch \
    python tools/gnn_batch_once.py ncu > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:aco_knn -s 4 -c 1 -o $O/${TAG}_k1_c5shard \
    python tools/k1_once.py c5shard > /dev/null 2>&1
ls -la $O | tail -20
