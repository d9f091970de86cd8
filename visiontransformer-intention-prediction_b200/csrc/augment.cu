// placeholder until the augmentation kernels land (next commit)
#include "bevgen_common.cuh"
extern "C" int bevgen_augment(const float*, float*, int, int, const bevgen_aug_params*, bevgen_stream_t) { return BEVGEN_E_UNSUPPORTED; }
