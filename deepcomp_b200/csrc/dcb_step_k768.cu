// Fused step kernel, CTA-size class <= 768 threads (80 registers per thread): see dcb_step_body.cuh
#define DCB_STEP_CLASS 768
#define DCB_STEP_REGS 80
#define DCB_STEP_KERNEL_NAME dcb_step_kernel_768
#define DCB_STEP_CLASS_FN(f) dcb_step_768_##f
#include "dcb_step_body.cuh"
