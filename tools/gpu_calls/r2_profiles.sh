# round 2 evidence: ncu launch lists (shares) + full captures of the dominant kernels.  Outputs: gpurun_out/r2p/
# tools/ncu_summary.py turns them into the tables under profiles/.
set -uo pipefail
O=gpurun_out/r2p; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_hifigan_b64.csv \
    python bench.py $BA > $O/ncu_hifigan.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bigvgan_b32.csv \
    python bench.py $BA --workload bigvgan_b32 > $O/ncu_bigvgan.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_vocos_huge_b128.csv \
    python bench.py $BA --workload vocos_huge_b128 > $O/ncu_vocos.log 2>&1
capture() {  # capture NAME KERNEL_REGEX SKIP COUNT BENCH_ARGS...
  local name="$1" rx="$2" skip="$3" cnt="$4"; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${rx}" -s "${skip}" -c "${cnt}" -f \
      -o "$O/prof_${name}" python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph \
      --steps 1 --warmup 3 "$@" > "$O/ncu_full_${name}.log" 2>&1
  ncu -i "$O/prof_${name}.ncu-rep" --page raw --csv > "$O/prof_${name}_raw.csv" 2>/dev/null
  rm -f "$O/prof_${name}.ncu-rep"
}
capture mrf_fused mrf_fused 44 11          # one forward: 9 C=128 pair launches + the C=64 and C=32 stages
capture conv_tc conv_tc 96 6               # C=256 stage convs
capture snake snake_aa 400 2 --workload bigvgan_b32
capture vocos_gemm conv_tc 171 2 --workload vocos_huge_b128
ls -la $O
