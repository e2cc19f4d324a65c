#!/usr/bin/env bash
# Round-end evidence run on the B200 box: benches for every workload (with the CPU baseline and the reference arm) + ncu
# launch lists + full captures of the dominant kernels.  Outputs land in gpurun_out/; tools/ncu_summary.py turns them
# into the tables under profiles/.
set -uo pipefail
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_hifigan_b64.log 2>&1
python bench.py --steps 10 --warmup 3 --workload bigvgan_b32 > gpurun_out/final_bench_bigvgan_b32.log 2>&1
python bench.py --steps 10 --warmup 3 --workload vocos_huge_b128 > gpurun_out/final_bench_vocos_huge_b128.log 2>&1
python bench.py --steps 20 --warmup 3 --workload hifigan_b1 --no-cpu-baseline > gpurun_out/final_bench_hifigan_b1.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.log 2>&1
# launch lists: skip the one-time weight packing (torch kernels of the first forward), keep ~5 forwards
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 260 --csv --log-file gpurun_out/launches_hifigan_b64.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 800 --csv --log-file gpurun_out/launches_bigvgan_b32.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --workload bigvgan_b32 > gpurun_out/ncu_bench_bigvgan.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 600 --csv --log-file gpurun_out/launches_vocos_huge_b128.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --workload vocos_huge_b128 > gpurun_out/ncu_bench_vocos.log 2>&1
# full captures: conv_tc in the C=128 stage of HiFiGAN (convs1, convs2, convs1), both fused MRF stages, one snake launch
# (a --set full report is ~35 MB and gpurun_out/ may carry 64 MiB: the raw page is exported on the box, the report dropped)
capture() {  # capture NAME KERNEL_REGEX SKIP COUNT BENCH_ARGS...
  local name="$1" rx="$2" skip="$3" cnt="$4"; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:${rx}" -s "${skip}" -c "${cnt}" -f \
      -o "gpurun_out/prof_${name}" python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline "$@" \
      > "gpurun_out/ncu_full_${name}.log" 2>&1
  ncu -i "gpurun_out/prof_${name}.ncu-rep" --page raw --csv > "gpurun_out/prof_${name}_raw.csv" 2>/dev/null
  rm -f "gpurun_out/prof_${name}.ncu-rep"
}
capture conv_tc conv_tc 105 3
capture mrf_fused mrf_fused 4 2
capture snake snake_aa 40 1 --workload bigvgan_b32
capture vocos_gemm conv_tc 171 2 --workload vocos_huge_b128
for f in gpurun_out/final_bench_*.log; do echo "== $f"; tail -c 400 "$f"; echo; done
