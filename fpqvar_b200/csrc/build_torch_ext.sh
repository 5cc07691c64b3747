#!/usr/bin/env bash
# Builds the torch extension over the C ABI: fpqvar_b200/dropin/quant_cuda<EXT_SUFFIX> (module name `quant_cuda`, the name of
# the reference's extension), linked against ../libfpq_b200.so (rpath $ORIGIN/..).  Host C++ only: the kernels are in the library.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$(dirname "$HERE")"
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_INC2 TORCH_LIB PY_INC EXT_SUFFIX <<<"$($PY - <<'PYEOF'
import sysconfig, logging
logging.disable(logging.CRITICAL)
from torch.utils import cpp_extension as ce
inc = ce.include_paths(); lib = ce.library_paths()
print(inc[0], inc[1], lib[0], sysconfig.get_paths()['include'], sysconfig.get_config_var('EXT_SUFFIX'))
PYEOF
)"
TARGET="$PKG/dropin/quant_cuda$EXT_SUFFIX"
SRC="$HERE/fpq_torch.cpp"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$SRC" ] && [ "$TARGET" -nt "$PKG/../include/fpq_b200.h" ]; then
  echo "build_torch_ext: $TARGET up to date"; exit 0
fi
[ -f "$PKG/libfpq_b200.so" ] || { echo "build_torch_ext: build libfpq_b200.so first (make -C fpqvar_b200/csrc)"; exit 1; }
mkdir -p "$HERE/build"
g++ -c "$SRC" -o "$HERE/build/fpq_torch.o" -DTORCH_EXTENSION_NAME=quant_cuda -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 \
    -I"$TORCH_INC" -I"$TORCH_INC2" -I"$PY_INC" -I/usr/local/cuda/include -O2 -std=c++17 -fPIC -w
g++ -shared "$HERE/build/fpq_torch.o" -o "$TARGET" -L"$PKG" -lfpq_b200 \
    -L"$TORCH_LIB" -L/usr/local/cuda/lib64 -lc10 -ltorch_cpu -ltorch -ltorch_python -lc10_cuda -ltorch_cuda -lcudart \
    -Wl,-rpath,'$ORIGIN/..' -Wl,-rpath,"$TORCH_LIB"
echo "build_torch_ext: wrote $TARGET"
