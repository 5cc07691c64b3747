// sign-split fake-quant kernels, kernel tie rule + the fpq_fake_quant_signsplit entry point
#define FPQ_SPLIT_TIE_PART 0
#include "fpq_split.inc.cuh"
