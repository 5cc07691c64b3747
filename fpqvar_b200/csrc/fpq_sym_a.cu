// symmetric fake-quant kernels, argmin tie rule (quantize_to_nearest_grid semantics)
#define FPQ_SYM_TIE_PART 1
#include "fpq_sym.inc.cuh"
