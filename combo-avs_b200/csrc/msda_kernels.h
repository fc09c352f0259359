// msda_kernels.h -- internal launcher interface between msda_api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "msda_common.cuh"

namespace msda {

struct Problem {
    int dtype, loc_dtype;
    int N, S, M, D, L, Lq, P;
    const void *value;
    const int64_t *shapes;  // device
    const int64_t *start;   // device
    const void *loc;
    const void *attn;
    bool have_host;    // host mirror of the geometry supplied
    LevelTable table;  // valid iff have_host
    int sm_count;
    int smem_optin;    // max dynamic shared memory per block (opt-in)
    int device = 0;    // CUDA ordinal the pointers live on (per-device one-time kernel attribute setup)
    // fused module prologue (ops/modules/ms_deform_attn.py:102-116): when `ref` is set, `loc` holds the raw sampling OFFSETS and
    // locations are ref + offset / (W_l, H_l); in the forward `attn` then holds the raw attention LOGITS (softmax over L*P is
    // taken in the kernel and written to `attn_out`), in the backward it holds the softmaxed weights saved by the forward.
    const void *ref = nullptr;   // [N, Lq, L, 2] (x, y), dtype of loc
    void *attn_out = nullptr;    // [N, Lq, M, L, P]
    const unsigned char *mask = nullptr;  // optional padding mask [N, S] (non-zero = padded pixel: its value row reads as zero)
};

struct BackwardIO {
    const void *grad_out;
    void *grad_value;
    void *grad_loc;
    void *grad_attn;
    void *workspace;
    size_t workspace_bytes;
};

void count_launch(int n = 1);

// generic: any D / L / P, f32 / f64 / bf16 (msda_generic.cu)
cudaError_t fwd_generic(const Problem &p, void *out, cudaStream_t st);
cudaError_t bwd_generic(const Problem &p, const BackwardIO &io, cudaStream_t st);
size_t bwd_generic_workspace(int dtype, int N, int S, int M, int D);
cudaError_t convert_f32_to_bf16(const float *src, void *dst, long long n, int sm_count, cudaStream_t st);

// warp32: D == 32, f32, warp-cooperative gather straight from L1/L2 (msda_warp32.cu)
bool warp32_supported(const Problem &p);
cudaError_t fwd_warp32(const Problem &p, void *out, cudaStream_t st);
cudaError_t bwd_warp32(const Problem &p, const BackwardIO &io, cudaStream_t st);

// slab32: D == 32, f32, host geometry: shared-memory slab gather (msda_slab32.cu)
bool slab32_supported(const Problem &p);
cudaError_t fwd_slab32(const Problem &p, void *out, cudaStream_t st);

// sorted32: D == 32, f32, host geometry: sort-and-segment backward (msda_sorted32.cu)
bool sorted32_supported(const Problem &p);
cudaError_t bwd_sorted32(const Problem &p, const BackwardIO &io, cudaStream_t st);

// msda_linear.cu: C[M, N] = A[M, K] . W[N, K]^T + bias on tcgen05 tensor cores (3xTF32 split: float32 accuracy)
bool linear_supported(int M, int N, int K);
bool linear_wgrad_supported(int M, int N, int K);
cudaError_t linear_tf32(const float *A, const float *W, const float *bias, float *C, int M, int N, int K, int relu, int split, int mn_major,
                        int device, int sm_count, int smem_optin, cudaStream_t st);

// msda_layernorm.cu: y = LayerNorm(a + b) forward / backward and column sums (bias gradients), HBM-bound single-pass kernels
bool layernorm_supported(int C);
size_t layernorm_backward_workspace(int M, int C, int sm_count);
cudaError_t add_layernorm_forward(const float *a, const float *b, const float *gamma, const float *beta, float *y, float *mean, float *rstd,
                                  int M, int C, float eps, int sm_count, cudaStream_t st);
cudaError_t layernorm_backward(const float *g, const float *a, const float *b, const float *gamma, const float *mean, const float *rstd, float *dx,
                               float *dgamma, float *dbeta, float *workspace, int M, int C, int sm_count, cudaStream_t st);
bool colsum_supported(int N);
size_t colsum_workspace(int M, int N, int sm_count);
cudaError_t colsum(const float *g, float *out, float *workspace, int M, int N, int sm_count, cudaStream_t st);

// zero the rows of grad[rows][row_elems] whose mask byte is set (masked_fill backward, ms_deform_attn.py:103-104)
cudaError_t zero_masked_rows(float *grad, const unsigned char *mask, long long rows, int row_elems, int sm_count, cudaStream_t st);
// softmax backward over the last L*P entries of every (frame, query, head) row, in place on grad (msda_generic.cu)
cudaError_t softmax_backward_rows(const float *attn, float *grad, long long rows, int LP, int sm_count, cudaStream_t st);

}  // namespace msda
