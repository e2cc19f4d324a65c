# round 2, call E: ncu launch lists restricted to the library's kernels (namespace fv::), full capture of fv_mrf_fused
set -uo pipefail
O=gpurun_out/r2e; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 2 --warmup 3"
for wl in hifigan_b64 bigvgan_b32 vocos_huge_b128; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fv:: -c 4000 --csv --log-file $O/launches_$wl.csv \
      python bench.py $BA --workload $wl > $O/ncu_$wl.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mrf_fused -s 4 -c 2 -f -o $O/prof_mrf_fused \
    python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 1 --warmup 3 > $O/ncu_full_mrf_fused.log 2>&1
ncu -i $O/prof_mrf_fused.ncu-rep --page raw --csv > $O/prof_mrf_fused_raw.csv 2>/dev/null
rm -f $O/prof_mrf_fused.ncu-rep
ls -la $O
