#!/usr/bin/env bash
# Builds libbimodal_b200.so (sm_100a only) next to the Python package. Usage: csrc/build.sh [-v]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libbimodal_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC)
if [[ "${1:-}" == "-v" ]]; then FLAGS+=(-Xptxas -v); fi
mkdir -p "$HERE/_obj"
pids=()
for f in pgd_step sampler scoring_ops attention_tc attention_split attention_occ engine comm phase_a grad_api; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE"/_obj/{pgd_step,sampler,scoring_ops,attention_tc,attention_split,attention_occ,engine,comm,phase_a,grad_api}.o -ldl
echo "built $OUT"
