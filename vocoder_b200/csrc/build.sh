#!/usr/bin/env bash
# Build libfv_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libfv_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC
       -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "${HERE}/_obj"
pids=()
for f in fv_api fv_simt fv_conv_tc fv_mrf_fused fv_frontend fv_debug; do
  ( "${NVCC}" "${FLAGS[@]}" -dc -c "${HERE}/${f}.cu" -o "${HERE}/_obj/${f}.o" > "${HERE}/_obj/${f}.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "${HERE}"/_obj/*.log; exit 1; fi
"${NVCC}" -gencode arch=compute_100a,code=sm_100a --shared -Xcompiler -fPIC -o "${OUT}" \
  "${HERE}/_obj/fv_api.o" "${HERE}/_obj/fv_simt.o" "${HERE}/_obj/fv_conv_tc.o" "${HERE}/_obj/fv_mrf_fused.o" "${HERE}/_obj/fv_frontend.o" "${HERE}/_obj/fv_debug.o"
echo "built ${OUT}"
