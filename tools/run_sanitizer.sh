#!/usr/bin/env bash
# compute-sanitizer over the small parity cases of every hand-written kernel family (SURVEY 5 "race detection" row):
# memcheck (out-of-bounds / misaligned shared and global accesses) and racecheck (shared-memory hazards between the
# generic-proxy writes of the epilogue warps and everything else).  Usage (GPU box): bash tools/run_sanitizer.sh [outdir]
# The selection covers fv_mrf_fused C = 16 / 32 / 64 (incl. the ALIAS staging configurations), conv_tc (single CTA,
# CTA pair, slab mainloop, 16-warp epilogue, strict operands), snake_aa (streaming + edge modes), dwconv_ln (cp.async pipeline), ISTFT, row-pair convs.
set -uo pipefail
OUT="${1:-gpurun_out/sanitizer}"
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
K='(test_mrf_fused_kernel and (16-45 or 32-52 or 64-384)) or test_conv1d_kernels or test_cuda_core_kernels or (test_generator_matches_reference_golden and (hifigan_small_stress or bigvgan_small_stress or vocos_small_stress)) or (test_snake_edge_modes and replicate-snakebeta-True) or (test_mrf_fused_pair_kernel_c128 and 3-1-300) or test_tiny_weights or (test_row_pair_conv_matches_the_plain_conv and 3-1-16)'
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1700 "$SAN" --tool "$tool" --print-limit 20 --error-exitcode 9 \
    python -m pytest tests/test_parity_gpu.py tests/test_round2_gpu.py -x -q -m gpu -k "$K" -p no:cacheprovider \
    > "$OUT/$tool.log" 2>&1
  echo "exit $? ($tool)" | tee -a "$OUT/$tool.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$OUT/$tool.log" | tail -n 8
done
