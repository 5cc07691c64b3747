// symmetric fake-quant kernels, kernel tie rule (quant_cuda.quant semantics) + the fpq_fake_quant entry point
#define FPQ_SYM_TIE_PART 0
#include "fpq_sym.inc.cuh"
