#!/bin/bash
# One gpurun command for a round's evidence (about 4 GPU-minutes):   gpurun --timeout 600 -- 'bash tools/profile_round.sh r02'
# Writes everything under gpurun_out/<tag>_*; summarise the .ncu-rep files here with tools/ncu_summary.py and copy what
# should be judged into profiles/.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
set -x
timeout 200 python -m pytest tests -m gpu -q > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log
timeout 120 python bench.py > $O/${TAG}_bench_1gpu.json 2> $O/${TAG}_bench.err
timeout 120 python tools/bench_configs.py > $O/${TAG}_bench_configs.txt 2>&1
timeout 60 python tools/bench_two_opt.py --variants v2 > $O/${TAG}_two_opt_bench.jsonl 2>&1
timeout 90 python tools/bench_gnn_train.py --iters 20 > $O/${TAG}_gnn_train_bench.jsonl 2>&1
# launch list of bench.py's step (kernel shares), then one full capture per kernel family
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aco_|tsp_|hadamard' -s 9 -c 30 --csv \
    --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:aco_knn -s 4 -c 1 -o $O/${TAG}_k1_knn \
    python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:tsp_update -s 4 -c 1 -o $O/${TAG}_k2_update \
    python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > /dev/null 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:two_opt_kernel -c 1 -o $O/${TAG}_k4_two_opt \
    python tools/bench_two_opt.py --sizes 500x256 --iters 1 --variants v2 > /dev/null 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:aco_list_kernel -s 1 -c 1 -o $O/${TAG}_cvrp_list \
    python tools/cvrp_once.py > /dev/null 2>&1
timeout 90 ncu --set full --clock-control none --import-source on -k regex:gnn_group_forward -s 2 -c 1 -o $O/${TAG}_gnn_group_c3 \
    python tools/gnn_eval_once.py C3 > /dev/null 2>&1
ls -la $O | tail -20
