# round 2, call Q: ncu --set full of the kernels changed late in the round (Snake ring look-ahead, row-pair convs of the C = 16 stage,
# conv_post with batched loads), final state
set -uo pipefail
O=gpurun_out/r2q; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 1 --warmup 3"
cap() {  # cap NAME WORKLOAD REGEX SKIP COUNT
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$3" -s "$4" -c "$5" -f -o $O/prof_$1 \
      python bench.py $BA --workload $2 > $O/ncu_$1.log 2>&1
  ncu -i $O/prof_$1.ncu-rep --page raw --csv > $O/prof_$1_raw.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
cap snake_c64 bigvgan_b32 snake_aa 400 2       # 4 forwards x 91 + 36: first Snake launches of the C = 64 stage
cap snake_c16 bigvgan_b32 snake_aa 436 2       # ... of the C = 16 stage
cap conv_c16 bigvgan_b32 conv_tc 461 6         # 4 forwards x 96 + 77: ups + first convs of the C = 16 stage (row-pair view)
cap conv_post hifigan_b64 conv_post 4 1
ls -la $O | head -20
