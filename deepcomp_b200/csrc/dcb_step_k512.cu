// Fused step kernel, CTA-size class <= 512 threads (128 registers per thread): see dcb_step_body.cuh
#define DCB_STEP_CLASS 512
#define DCB_STEP_REGS 128
#define DCB_STEP_KERNEL_NAME dcb_step_kernel_512
#define DCB_STEP_CLASS_FN(f) dcb_step_512_##f
#include "dcb_step_body.cuh"
