#!/bin/bash
# Run on the GPU box (gpurun): launch list + full ncu capture of the dominant kernel of bench.py's step.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aco_|tsp_|hadamard' -s 9 -c 30 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > gpurun_out/r01_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aco_knn -s 4 -c 1 -o gpurun_out/r01_k1_knn \
    python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > gpurun_out/r01_k1_bench.log 2>&1
ncu --set full --clock-control none -k regex:tsp_update -s 4 -c 1 -o gpurun_out/r01_k2_update \
    python bench.py --steps 3 --warmup 3 --colonies 256 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
