#!/usr/bin/env bash
# TEST INFRASTRUCTURE (checker only).  Builds the reference's OWN CUDA extensions
# (raymarching/src, gridencoder/src) for sm_100a from the sources where they lie
# under /root/reference, into oracle/_ref/ (git-ignored, travels with gpurun).
# No reference source is copied into this repo.  The only deviation from the
# reference's build flags (raymarching/backend.py:6-9, gridencoder/backend.py:6-9)
# is -std=c++17 (torch 2.11 headers reject C++14) and an explicit sm_100a gencode.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
PY=${PYTHON:-python}
TORCH_DIR=$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))' 2>/dev/null)
PYINC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
COMMON="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a \
 -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__ \
 --expt-relaxed-constexpr -Xcompiler -fPIC -D_GLIBCXX_USE_CXX11_ABI=1 \
 -I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PYINC -w"
LINK="-shared -L$TORCH_DIR/lib -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python -Xlinker -rpath -Xlinker $TORCH_DIR/lib"

build_one () {  # $1 = package dir in the reference, $2 = module name
  local pkg=$1 mod=$2
  if [ -f "$OUT/$mod.so" ] && [ "${FORCE:-0}" != "1" ]; then echo "[build_ref] $mod.so present"; return; fi
  echo "[build_ref] compiling $pkg -> $OUT/$mod.so (several minutes)"
  $NVCC $COMMON -DTORCH_EXTENSION_NAME=$mod -c "$REF/$pkg/src/$pkg.cu" -o "$OUT/$mod.cu.o" &
  local p1=$!
  g++ -O3 -std=c++17 -fPIC -D_GLIBCXX_USE_CXX11_ABI=1 -DTORCH_EXTENSION_NAME=$mod \
     -I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PYINC -I/usr/local/cuda/include \
     -c "$REF/$pkg/src/bindings.cpp" -o "$OUT/$mod.bind.o" -w
  wait $p1
  $NVCC $LINK "$OUT/$mod.cu.o" "$OUT/$mod.bind.o" -o "$OUT/$mod.so"
  rm -f "$OUT/$mod.cu.o" "$OUT/$mod.bind.o"
  echo "[build_ref] built $OUT/$mod.so"
}
build_one raymarching _raymarching_ref &
build_one gridencoder _gridencoder_ref &
wait
ls -la "$OUT"
