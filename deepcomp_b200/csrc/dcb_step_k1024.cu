// Fused step kernel, CTA-size class <= 1024 threads (64 registers per thread): see dcb_step_body.cuh
#define DCB_STEP_CLASS 1024
#define DCB_STEP_REGS 64
#define DCB_STEP_KERNEL_NAME dcb_step_kernel_1024
#define DCB_STEP_CLASS_FN(f) dcb_step_1024_##f
#include "dcb_step_body.cuh"
