/*
 * msda_b200.h -- C ABI of libmsda_b200.so: multi-scale deformable attention core for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for COMBO-AVS's one native extension.  Each entry point replaces one symbol
 * of the reference's pybind11 module `MultiScaleDeformableAttention`
 * (/root/reference/models/modeling/pixel_decoder/ops/src/vision.cpp:18-21):
 *
 *   msda_forward   <->  ms_deform_attn_forward   (src/ms_deform_attn.h:25-44  -> src/cuda/ms_deform_attn_cuda.cu:25-85)
 *   msda_backward  <->  ms_deform_attn_backward  (src/ms_deform_attn.h:46-66  -> src/cuda/ms_deform_attn_cuda.cu:88-158)
 *
 * Differences from the reference boundary, all deliberate:
 *   - plain pointers and sizes, no torch/ATen types: the caller allocates every output (torch.empty is enough;
 *     the library zero-fills whatever the selected kernel needs) and passes the device ordinal and the
 *     cudaStream_t to run on.  The reference allocates with at::zeros and uses the ambient current stream.
 *   - every function returns a status (0 ok, <0 argument error, >0 cudaError_t) and records a message readable
 *     with msda_last_error(); the reference only printf()s launch errors (ms_deform_im2col_cuda.cuh:953-957,1326-1330).
 *   - bf16 is accepted in addition to the reference's float/double dispatch (ms_deform_attn_cuda.cu:69,139).
 *   - there is no CPU implementation and no fallback (the reference's CPU entry points are AT_ERROR stubs,
 *     src/cpu/ms_deform_attn_cpu.cpp:31,44).
 *
 * Tensor layouts are the reference's (row-major, contiguous; ms_deform_attn_cuda.cu:45-53):
 *   value          [N, S, M, D]            dtype
 *   spatial_shapes [L, 2]  (H_l, W_l)      int64, device memory          (host mirror optional, see below)
 *   level_start    [L]                     int64, device memory
 *   sampling_loc   [N, Lq, M, L, P, 2]     (x, y) normalised to [0,1]; loc_dtype
 *   attn_weight    [N, Lq, M, L, P]        loc_dtype
 *   output / grad_output [N, Lq, M*D]      dtype
 *   grad_value like value; grad_loc / grad_attn like sampling_loc / attn_weight.
 *
 * Thread safety: re-entrant.  The only process-wide state is the variant override (msda_select_variant, meant for
 * profiling A/B runs) and one-time cudaFuncSetAttribute calls.  msda_last_error() is thread-local.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_VERSION 102 /* major*100 + minor: 1.0 the op, 1.1 + msda_linear_*, 1.2 + msda_add_layernorm_* / msda_colsum */

/* element types */
#define MSDA_F32 0
#define MSDA_BF16 1
#define MSDA_F64 2

/* status codes (< 0: argument errors; > 0: cudaError_t values) */
#define MSDA_OK 0
#define MSDA_ERR_NULL -1      /* a required pointer is NULL */
#define MSDA_ERR_DIM -2       /* a dimension is <= 0 or overflows int32 indexing */
#define MSDA_ERR_DTYPE -3     /* unsupported dtype / loc_dtype combination */
#define MSDA_ERR_WORKSPACE -4 /* workspace too small (see msda_backward_workspace_bytes) */
#define MSDA_ERR_VARIANT -5   /* unknown variant name, or the forced variant cannot run this shape */
#define MSDA_ERR_STEP -6      /* N % min(N, im2col_step) != 0 (ms_deform_attn_cuda.cu:55-57) */

/*
 * Forward.  out[n,q,m*D+c] = sum_{l,p} attn * bilinear(value level l, loc*(W,H) - 0.5), zero outside the map
 * (ms_deform_im2col_cuda.cuh:242-304, 38-89).
 *
 * shapes_host / level_start_host: optional host copies of the two int64 arrays (NULL if unknown).  With them the
 * library can pick shape-specialised kernels (shared-memory slabs) without a device->host sync; without them
 * only the kernels that read the geometry from device memory are used.  Contents must equal the device arrays.
 * im2col_step: accepted for API compatibility; validated like the reference, otherwise unused (one launch
 * covers the whole batch).
 */
int msda_forward(int dtype, int loc_dtype,
                 const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const void *sampling_loc, const void *attn_weight, void *output,
                 int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                 const int64_t *shapes_host, const int64_t *level_start_host,
                 int device, void *stream);

/*
 * Backward (ms_deform_im2col_cuda.cuh:306-408, 92-164).  Writes all three gradients in full; the caller does
 * not need to zero them.  workspace: device scratch of at least msda_backward_workspace_bytes(...) bytes
 * (may be NULL when that is 0).
 */
int msda_backward(int dtype, int loc_dtype,
                  const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                  const void *sampling_loc, const void *attn_weight, const void *grad_output,
                  void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                  int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                  const int64_t *shapes_host, const int64_t *level_start_host,
                  void *workspace, size_t workspace_bytes,
                  int device, void *stream);

/* Scratch bytes msda_backward needs for this problem: the fp32 accumulator of grad_value for bf16 values, else 0. */
size_t msda_backward_workspace_bytes(int dtype, int loc_dtype, int N, int S, int M, int D, int L, int Lq, int P);

/*
 * Force a kernel variant for subsequent calls (profiling A/B runs): pass_ 0 = forward, 1 = backward;
 * name = "auto" restores the default choice.  Returns MSDA_ERR_VARIANT for unknown names.
 * msda_variant_name writes the variant the library would run (or last ran) for the pass.
 */
int msda_select_variant(int pass_, const char *name);
const char *msda_variant_names(int pass_); /* comma-separated list */
const char *msda_last_variant(int pass_);  /* name of the variant last launched for the pass (process-wide) */

/* Number of kernel launches (memsets excluded) issued by this library since load, for gpu_launches accounting. */
uint64_t msda_launch_count(void);

const char *msda_last_error(void); /* thread-local, never NULL */
int msda_version(void);

/*
 * Fused module prologue.  The reference's MSDeformAttn.forward computes, between its Linear
 * layers and the op (ops/modules/ms_deform_attn.py:102-116):
 *     attention_weights  = softmax(attention_logits.view(N, Lq, M, L*P), -1)
 *     sampling_locations = reference_points[:, :, None, :, None, :] + sampling_offsets / (W_l, H_l)      (2-d reference points)
 * as separate elementwise kernels that write and re-read 31.6 MB of locations and 15.8 MB of weights per layer.  These two
 * entry points take the Linear outputs directly: the forward stages offsets / logits / reference points of a pass in shared
 * memory, transforms them in place (ex2-based exp and one reciprocal per row for the softmax, one FMA with 1/W, 1/H per
 * coordinate: a few ulp from the reference's expf / division kernels, far inside its 1e-5 tolerance) and writes the softmaxed weights
 * once (attention_weights_out, needed by the backward); the backward rebuilds the locations while it sorts, writes
 * d/d(sampling_offsets) directly (the (W, H) factors of ms_deform_im2col_cuda.cuh:126-127 cancel) and finishes with an
 * in-place softmax backward, so grad_attention_logits is d/d(logits).  reference_points: [N, Lq, L, 2] (x, y), no gradient.
 * padding_mask (optional, NULL = none): [N, S] bytes, non-zero = padded pixel.  It replaces the module's
 * value.masked_fill(mask, 0) (ms_deform_attn.py:103-104) and its backward: masked pixels read as zero rows while the slab is
 * staged, and their grad_value rows are zeroed at the end, so `value` is passed unmasked.
 * float32, D = 32, host geometry required; msda_fused_supported() tells whether a shape qualifies (else compose the unfused
 * entry points as the reference does).  Everything else (layouts, statuses, stream / device) as msda_forward / msda_backward.
 */
int msda_fused_supported(int dtype, int N, int S, int M, int D, int L, int Lq, int P,
                         const int64_t *shapes_host, const int64_t *level_start_host, int device);
int msda_fused_forward(int dtype,
                       const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *reference_points, const void *sampling_offsets, const void *attention_logits,
                       const unsigned char *padding_mask, void *output, void *attention_weights_out,
                       int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                       const int64_t *shapes_host, const int64_t *level_start_host, int device, void *stream);
int msda_fused_backward(int dtype,
                        const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *reference_points, const void *sampling_offsets, const void *attention_weights,
                        const unsigned char *padding_mask,
                        const void *grad_output, void *grad_value, void *grad_sampling_offsets, void *grad_attention_logits,
                        int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                        const int64_t *shapes_host, const int64_t *level_start_host, int device, void *stream);

/*
 * Host-buffer entry point: forward + backward of the core on tensors in HOST memory (pinned memory makes the copies
 * asynchronous).  The batch is cut into chunks of `chunk_frames` frames that flow through three streams owned by the
 * pipeline object (host->device copies, kernels, device->host copies) with `slots` (2..4) device staging buffers in
 * rotation, so PCIe runs full duplex and the kernels hide under the copies.  Semantics per frame are exactly
 * msda_forward followed by msda_backward (same reference lines); the host geometry arrays are required.  The call
 * returns when all four results are in host memory.  h2d_bytes / d2h_bytes (optional) receive the bytes moved.
 * A pipeline object serves one call at a time (not re-entrant); create one per thread / device.
 */
typedef struct msda_host_pipeline msda_host_pipeline;
int msda_host_pipeline_create(msda_host_pipeline **out, int device, int chunk_frames, int slots);
int msda_host_pipeline_destroy(msda_host_pipeline *p);
const char *msda_host_pipeline_error(const msda_host_pipeline *p);
int msda_forward_backward_host(msda_host_pipeline *p, int dtype, int loc_dtype,
                               const void *value, const int64_t *shapes_host, const int64_t *level_start_host,
                               const void *sampling_loc, const void *attn_weight, const void *grad_output,
                               void *output, void *grad_value, void *grad_sampling_loc, void *grad_attn_weight,
                               int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                               uint64_t *h2d_bytes, uint64_t *d2h_bytes);

/*
 * Row f2: the Linear layers around the op in the deformable encoder layer -- reference
 * models/modeling/pixel_decoder/msdeformattn.py:119-134 (linear1 / linear2 of the FFN) and ops/modules/ms_deform_attn.py:105-128
 * (value_proj, sampling_offsets, attention_weights, output_proj), which the reference runs as strict-fp32 cuBLAS GEMMs.
 *
 *   y[M, N] = x[M, K] . weight[N, K]^T + bias[N]            (torch.nn.functional.linear; row-major, float32, 16-byte aligned)
 *
 * on the tcgen05 tensor cores with every operand split into two TF32 halves (3 MMAs per product: float32-level accuracy).
 * Needs K % 4 == 0 and N % 16 == 0 (msda_linear_supported).  The same entry point serves dgrad: pass grad_y as x, weight^T
 * ([K, N] row-major) as weight, NULL bias -- and wgrad: grad_weight[N, K] = linear(grad_y^T [N, M], x^T [K, M]), where the long
 * reduction over M is split over the SMs (partial tiles are added into a zeroed result) -- or, without materialising the
 * transposes, msda_linear_wgrad below.  flags: MSDA_LINEAR_RELU fuses max(., 0); MSDA_LINEAR_TF32_ONLY skips the split
 * (plain TF32 accuracy, profiling only).
 */
#define MSDA_LINEAR_RELU 1
#define MSDA_LINEAR_TF32_ONLY 2
int msda_linear_supported(int M, int N, int K);
int msda_linear_forward(const void *x, const void *weight, const void *bias, void *y, int M, int N, int K, int flags, int device, void *stream);
/* grad_weight[N, K] = grad_y[M, N]^T . x[M, K] (the weight gradient of the same layer), both operands read as they lie in memory
 * (transposed-operand tcgen05.mma), the contraction over the M rows split over the SMs.  Needs N % 4 == 0 and K % 32 == 0. */
int msda_linear_wgrad_supported(int M, int N, int K);
int msda_linear_wgrad(const void *grad_y, const void *x, void *grad_weight, int M, int N, int K, int device, void *stream);


/*
 * Row f2, the elementwise epilogue of the encoder layer's GEMMs -- reference models/modeling/pixel_decoder/msdeformattn.py:119-134:
 *     src = norm1(src + dropout1(self_attn(...)));   src = norm2(src + dropout2(linear2(...)))        (dropout 0 in COMBO's configs)
 * i.e. y[M, C] = LayerNorm(a + b) * gamma + beta (torch.nn.LayerNorm semantics: biased variance, eps inside the square root), float32,
 * row-major, 16-byte aligned, C a multiple of 128 up to 512 (msda_layernorm_supported).  `b` may be NULL (plain LayerNorm).  The forward
 * saves mean[M] and rstd[M]; the backward recomputes the normalised row from a + b and returns grad_x (the gradient of BOTH a and b),
 * grad_gamma[C] and grad_beta[C], reduced over the rows in a fixed order (deterministic; scratch of
 * msda_layernorm_backward_workspace_bytes(M, C, device) bytes from the caller -- the library allocates nothing).
 * msda_colsum: out[N] = sum over the M rows of g[M, N] -- the bias gradient of a Linear layer (torch: grad_y.sum(0)); N % 4 == 0, N <= 4096.
 */
int msda_layernorm_supported(int C);
size_t msda_layernorm_backward_workspace_bytes(int M, int C, int device);
int msda_add_layernorm_forward(const void *a, const void *b, const void *gamma, const void *beta, void *y, void *mean, void *rstd,
                               int M, int C, float eps, int device, void *stream);
int msda_layernorm_backward(const void *grad_y, const void *a, const void *b, const void *gamma, const void *mean, const void *rstd,
                            void *grad_x, void *grad_gamma, void *grad_beta, void *workspace, size_t workspace_bytes,
                            int M, int C, int device, void *stream);
int msda_colsum_supported(int N);
size_t msda_colsum_workspace_bytes(int M, int N, int device);
int msda_colsum(const void *g, void *out, void *workspace, size_t workspace_bytes, int M, int N, int device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
