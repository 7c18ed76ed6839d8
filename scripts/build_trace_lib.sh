#!/bin/bash
# Builds the -DRSR_TRACE variant of the library (in-kernel phase timing of the recurrence kernels) next to the
# product library: rsrgan_b200/librsrgan_trace.so.  Use with RSR_LIB=rsrgan_b200/librsrgan_trace.so.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
C="$HERE/../rsrgan_b200/csrc"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static -DRSR_TRACE)
mkdir -p "$C/build/trace"
pids=()
for f in gemm_sm100 elementwise batchnorm lstmp_sm100 lstmp_cluster_sm100 lstmp_pair_sm100 peer_allreduce; do
  "$NVCC" "${FLAGS[@]}" -c "$C/$f.cu" -o "$C/build/trace/$f.o" > "$C/build/trace/$f.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$C"/build/trace/*.log; exit 1; fi
"$NVCC" -shared -cudart static -o "$HERE/../rsrgan_b200/librsrgan_trace.so" "$C"/build/trace/*.o
echo "built librsrgan_trace.so"
