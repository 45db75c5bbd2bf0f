// Fused step kernel: host-side dispatch to the CTA-size classes (dcb_step_k<threads>.cu, each an instance of
// dcb_step_body.cuh compiled with its own register budget).
#include "dcb_internal.h"

#define DCB_DECLARE_CLASS(c)                                                                            \
    cudaError_t dcb_step_##c##_set_smem(int n_bs, size_t smem);                                         \
    int dcb_step_##c##_regs(int n_bs);                                                                  \
    cudaError_t dcb_step_##c##_launch(const StepArgs &a, int threads, int grid, size_t smem, cudaStream_t s);
DCB_DECLARE_CLASS(256)
DCB_DECLARE_CLASS(512)
DCB_DECLARE_CLASS(704)
DCB_DECLARE_CLASS(768)
DCB_DECLARE_CLASS(1024)

#define DCB_BY_CLASS(threads, f, ...)                                            \
    ((threads) <= 256 ? dcb_step_256_##f(__VA_ARGS__)                            \
     : (threads) <= 512 ? dcb_step_512_##f(__VA_ARGS__)                          \
     : (threads) <= 704 ? dcb_step_704_##f(__VA_ARGS__)                          \
     : (threads) <= 768 ? dcb_step_768_##f(__VA_ARGS__) : dcb_step_1024_##f(__VA_ARGS__))

size_t dcb_step_smem_bytes(int kind, int N, int M, int E, int var) { return (size_t)dcb_smem_layout(kind, N, M, E, var).total; }

cudaError_t dcb_step_set_smem_limit(int threads, int n_bs, size_t smem) { return DCB_BY_CLASS(threads, set_smem, n_bs, smem); }

int dcb_step_regs_per_thread(int threads, int n_bs) { return DCB_BY_CLASS(threads, regs, n_bs); }

cudaError_t dcb_launch_step(const StepArgs &a, int threads, int grid, size_t smem, cudaStream_t s) {
    return DCB_BY_CLASS(threads, launch, a, threads, grid, smem, s);
}
