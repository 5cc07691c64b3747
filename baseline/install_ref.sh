#!/usr/bin/env bash
# MEASUREMENT INFRASTRUCTURE ONLY -- never on the product path.
#
# "Installs" the UNMODIFIED reference (PKU-SEC-Lab/FPQVAR, /root/reference) under baseline/_ref/FPQVAR so that
# bench.py's reference legs can run the reference's OWN code on the GPU box (where /root/reference does not exist):
#   * its torch CPU path        models_fp_quant*/quant_utils.py::fp_quant_*_per_group            (cpu_baseline, --impl reference)
#   * its GPU path              the same files' fp_quant_*_cuda around its quant_cuda extension  (reference_gpu_path)
#   * its model                 models_fp_quant_transform_rotate.build_vae_var + VAR.autoregressive_infer_cfg,
#                               rotate_model / transform_model / quantize_VAR and the shipped GALT factors
#                               learnable_transformation/best_lambda_var30/*.pt                  (generation_reference_model)
# The reference is a flat research repository without package metadata, so `pip install` does not apply; this copies
# the tree (minus figures and build debris) and builds its extension under its own module name `quant_cuda`
# (quant/quant.cpp:31, PYBIND11_MODULE(TORCH_EXTENSION_NAME)) from the sources where they lie, reusing the kernel object
# that oracle/build_ref.sh compiled.  baseline/_ref/ is git-ignored (no reference source enters the history) but NOT
# gpurun-ignored: it travels to the GPU box like the built .so files.
set -euo pipefail
REF=${FPQ_REFERENCE_ROOT:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_ref"
if [ ! -d "$REF/models_fp_quant" ]; then
  echo "install_ref: $REF not present -- skipping (an earlier install in $OUT is used as it is)"
  exit 0
fi
mkdir -p "$OUT/FPQVAR"
# the tree, unmodified
(cd "$REF" && tar -cf - --exclude=readme_figs --exclude=quant/build --exclude='*.egg-info' --exclude=.git --exclude=__pycache__ .) | (cd "$OUT/FPQVAR" && tar -xf -)
chmod -R u+w "$OUT/FPQVAR"
# the extension under its own name
bash "$ROOT/oracle/build_ref.sh"
KOBJ="$ROOT/oracle/_ref/obj/quant_kernel.o"
[ -f "$KOBJ" ] || { echo "install_ref: $KOBJ missing (oracle/build_ref.sh failed?)"; exit 1; }
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_INC2 TORCH_LIB PY_INC EXT_SUFFIX <<<"$($PY - <<'EOF'
import sysconfig, logging
logging.disable(logging.CRITICAL)
from torch.utils import cpp_extension as ce
inc = ce.include_paths(); lib = ce.library_paths()
print(inc[0], inc[1], lib[0], sysconfig.get_paths()['include'], sysconfig.get_config_var('EXT_SUFFIX'))
EOF
)"
TARGET="$OUT/FPQVAR/quant_cuda$EXT_SUFFIX"
if [ ! -f "$TARGET" ] || [ "$KOBJ" -nt "$TARGET" ]; then
  mkdir -p "$OUT/obj"
  g++ -c "$REF/quant/quant.cpp" -o "$OUT/obj/quant.o" -DTORCH_EXTENSION_NAME=quant_cuda -DTORCH_API_INCLUDE_EXTENSION_H \
      -D_GLIBCXX_USE_CXX11_ABI=1 -I"$TORCH_INC" -I"$TORCH_INC2" -I"$PY_INC" -I/usr/local/cuda/include -O2 -std=c++17 -fPIC -w
  g++ -shared "$OUT/obj/quant.o" "$KOBJ" -o "$TARGET" \
      -L"$TORCH_LIB" -L/usr/local/cuda/lib64 -lc10 -ltorch_cpu -ltorch -ltorch_python -lc10_cuda -ltorch_cuda -lcudart \
      -Wl,-rpath,"$TORCH_LIB"
fi
echo "install_ref: reference tree in $OUT/FPQVAR, extension $TARGET"
