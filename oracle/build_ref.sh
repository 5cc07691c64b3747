#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never on the product path.
#
# Compiles the UNMODIFIED reference extension (quant/quant.cpp + quant/quant_kernel.cu)
# from the sources where they lie under /root/reference into oracle/_ref/ as a Python
# module named `ref_quant_cuda` (renamed via -DTORCH_EXTENSION_NAME so it cannot shadow
# this repo's own drop-in `quant_cuda`).  No reference source is copied into the repo;
# only the built .so lands in oracle/_ref/ (git-ignored, but it travels with gpurun).
#
# The reference's own build system (quant/setup.py) is not used: it would write into the
# read-only source tree.  This is the equivalent two-compile + link recipe.
set -euo pipefail
REF=${FPQ_REFERENCE_ROOT:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/quant/quant_kernel.cu" ]; then
  echo "build_ref: $REF/quant not present -- skipping (prebuilt files in $OUT are used as-is)"
  exit 0
fi
mkdir -p "$OUT/obj"
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_INC2 TORCH_LIB PY_INC EXT_SUFFIX <<<"$($PY - <<'EOF'
import sysconfig, warnings, logging
logging.disable(logging.CRITICAL)
from torch.utils import cpp_extension as ce
inc = ce.include_paths(); lib = ce.library_paths()
print(inc[0], inc[1], lib[0], sysconfig.get_paths()['include'], sysconfig.get_config_var('EXT_SUFFIX'))
EOF
)"
TARGET="$OUT/ref_quant_cuda$EXT_SUFFIX"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$REF/quant/quant_kernel.cu" ] && [ "$TARGET" -nt "$REF/quant/quant.cpp" ]; then
  echo "build_ref: $TARGET up to date"; exit 0
fi
COMMON="-DTORCH_EXTENSION_NAME=ref_quant_cuda -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 -I$TORCH_INC -I$TORCH_INC2 -I$PY_INC -I/usr/local/cuda/include"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
echo "build_ref: nvcc quant_kernel.cu (sm_100a; several minutes)"
$NVCC -c "$REF/quant/quant_kernel.cu" -o "$OUT/obj/quant_kernel.o" $COMMON \
  -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --compiler-options -fPIC \
  -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__ \
  --expt-relaxed-constexpr -w &
P1=$!
echo "build_ref: g++ quant.cpp"
g++ -c "$REF/quant/quant.cpp" -o "$OUT/obj/quant.o" $COMMON -O2 -std=c++17 -fPIC -w
wait $P1
g++ -shared "$OUT/obj/quant.o" "$OUT/obj/quant_kernel.o" -o "$TARGET" \
  -L"$TORCH_LIB" -L/usr/local/cuda/lib64 -lc10 -ltorch_cpu -ltorch -ltorch_python -lc10_cuda -ltorch_cuda -lcudart \
  -Wl,-rpath,"$TORCH_LIB"
echo "build_ref: wrote $TARGET"
