# round 2, call F: launch lists (fixed kernel filter) + full captures of the BigVGAN small-C convs
set -uo pipefail
O=gpurun_out/r2f; mkdir -p $O
RX='regex:conv_tc|mrf_fused|snake_aa|dwconv_ln|conv_post|pack_input|istft_ola|act_cast|resample|noise_conv|unpack_output|conv_simt'
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 2 --warmup 3"
for wl in hifigan_b64 bigvgan_b32 vocos_huge_b128; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 4000 --csv --log-file $O/launches_$wl.csv \
      python bench.py $BA --workload $wl > $O/ncu_$wl.log 2>&1
done
capture() {  # capture NAME SKIP COUNT
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s "$2" -c "$3" -f -o $O/prof_$1 \
      python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 1 --warmup 3 \
      --workload bigvgan_b32 > $O/ncu_full_$1.log 2>&1
  ncu -i $O/prof_$1.ncu-rep --page raw --csv > $O/prof_$1_raw.csv 2>/dev/null
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1_source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
capture bigvgan_c64 423 6     # stage C=64: ups + first convs (96 conv_tc launches per forward, 4 forwards skipped)
capture bigvgan_c32 442 6     # stage C=32
capture bigvgan_c16 461 6     # stage C=16
ls -la $O
