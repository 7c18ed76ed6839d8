#!/bin/bash
# Builds librsrgan_sm100.so (the C-ABI library) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../librsrgan_sm100.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v -cudart static ${RSR_EXTRA_NVCC_FLAGS:-})
mkdir -p "$HERE/build"
pids=()
for f in gemm_sm100 elementwise batchnorm lstmp_sm100 lstmp_cluster_sm100 lstmp_pair_sm100 peer_allreduce; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" > "$HERE/build/$f.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$HERE"/build/*.log; exit 1; fi
"$NVCC" -shared -cudart static -o "$OUT" "$HERE"/build/gemm_sm100.o "$HERE"/build/elementwise.o "$HERE"/build/batchnorm.o "$HERE"/build/lstmp_sm100.o "$HERE"/build/lstmp_cluster_sm100.o "$HERE"/build/lstmp_pair_sm100.o "$HERE"/build/peer_allreduce.o
echo "built $OUT"
