#!/usr/bin/env bash
# Round-end evidence run on the B200 box: benches for every workload + ncu launch lists + full captures.
set -uo pipefail
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_hifigan_b64.log 2>&1
python bench.py --steps 10 --warmup 3 --workload bigvgan_b32 > gpurun_out/final_bench_bigvgan_b32.log 2>&1
python bench.py --steps 10 --warmup 3 --workload vocos_huge_b128 > gpurun_out/final_bench_vocos_huge_b128.log 2>&1
python bench.py --steps 20 --warmup 3 --workload hifigan_b1 > gpurun_out/final_bench_hifigan_b1.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 420 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 760 --csv --log-file gpurun_out/launches_bigvgan_r01.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --workload bigvgan_b32 > gpurun_out/ncu_bench_bigvgan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 28 -c 3 -f -o gpurun_out/prof_conv_r01 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:snake_aa -s 30 -c 2 -f -o gpurun_out/prof_snake_r01 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --workload bigvgan_b32 > gpurun_out/ncu_full_snake.log 2>&1
tail -c 300 gpurun_out/final_bench_hifigan_b64.log
