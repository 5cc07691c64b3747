// sign-split fake-quant kernels, argmin tie rule
#define FPQ_SPLIT_TIE_PART 1
#include "fpq_split.inc.cuh"
