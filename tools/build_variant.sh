#!/bin/bash
# Build an experimental variant of libsmartcore_kmeans_cuda.so for A/B measurements on the GPU box:
#   tools/build_variant.sh NAME "-DSCKM_EXPERIMENT_X ..."   ->  smartcore_b200/lib/libsmartcore_kmeans_cuda.NAME.so
# Select it at run time with SCKM_LIB_VARIANT=NAME (smartcore_b200/cabi.py).  Objects go to build/variants/NAME (git-ignored).
set -e
NAME="$1"; FLAGS="$2"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="$ROOT/smartcore_b200/csrc"; OBJ="$ROOT/build/variants/$NAME"
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
pids=()
for f in "$SRC"/*.cu; do
  b=$(basename "$f" .cu)
  $NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $FLAGS -c "$f" -o "$OBJ/$b.o" 2> "$OBJ/$b.ptxas.log" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p || { cat "$OBJ"/*.ptxas.log | grep -i error; exit 1; }; done
$NVCC $ARCH -shared -o "$ROOT/smartcore_b200/lib/libsmartcore_kmeans_cuda.$NAME.so" "$OBJ"/*.o -ldl -lpthread
echo "built smartcore_b200/lib/libsmartcore_kmeans_cuda.$NAME.so"
