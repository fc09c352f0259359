// msda_generic.cu -- shape-generic MSDeformAttn kernels: any head width D, any level / point count,
// float32 / float64 / bfloat16.  This is the fallback family; it covers what the reference's seven backward
// kernel templates cover between them (ops/src/cuda/ms_deform_im2col_cuda.cuh:306-925 and the D-switch at
// :961-1331) with ONE design: a warp owns one (frame, query, head) and strides its lanes over channels, so
//   - grad_sampling_loc / grad_attn_weight are plain warp-shuffle reductions (no shared-memory cache, no serial
//     thread-0 loop, no global atomics whatever D is),
//   - the geometry and the per-sample scalars are computed once per warp iteration, not once per channel.
// grad_value uses one atomicAdd per (neighbour, channel) like the reference; bf16 accumulates in an fp32
// workspace that is converted at the end (bf16 atomics would lose most of the sum).
#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// out[n,q,m,c]: warp per (n,q,m,32-channel chunk)
template <typename TV, typename TM>
__global__ void __launch_bounds__(kThreads)
fwd_generic_kernel(const TV *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
                   const TM *__restrict__ loc, const TM *__restrict__ attn, TV *__restrict__ out,
                   int S, int M, int D, int L, int Lq, int P, int chunks, long long n_units)
{
    using CT = typename Compute<TV>::type;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long stride = (long long)gridDim.x * kWarps;
    const long long pix = (long long)M * D;
    for (long long u = warp0; u < n_units; u += stride) {
        const int chunk = (int)(u % chunks);
        const long long qm = u / chunks;          // (n*Lq + q)*M + m
        const int m = (int)(qm % M);
        const long long n = qm / ((long long)M * Lq);
        const int c = chunk * 32 + lane;
        const bool live = c < D;
        const TV *vhead = value + n * S * pix + (long long)m * D + (live ? c : 0);
        CT acc = 0;
        long long s = qm * L * P;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const TV *vl = vhead + start[l] * pix;
            for (int p = 0; p < P; ++p, ++s) {
                Taps<CT> t;
                if (!make_taps<CT>(to_compute(loc[2 * s]), to_compute(loc[2 * s + 1]), H, W, t)) continue;
                const CT a = to_compute(attn[s]);
                const TV *r0 = vl + ((long long)t.y0 * W + t.x0) * pix;
                const TV *r1 = r0 + (long long)W * pix;
                CT v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                if (live) {
                    if (t.top && t.lef) v1 = to_compute(r0[0]);
                    if (t.top && t.rig) v2 = to_compute(r0[pix]);
                    if (t.bot && t.lef) v3 = to_compute(r1[0]);
                    if (t.bot && t.rig) v4 = to_compute(r1[pix]);
                }
                acc += (t.hh * t.hw * v1 + t.hh * t.lw * v2 + t.lh * t.hw * v3 + t.lh * t.lw * v4) * a;
            }
        }
        if (live) out[qm * D + c] = from_compute<TV, CT>(acc);
    }
}

// warp per (n,q,m); lanes stride over channels; GA = accumulation type of grad_value (float for f32/bf16).
template <typename TV, typename TM, typename GA>
__global__ void __launch_bounds__(kThreads)
bwd_generic_kernel(const TV *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
                   const TM *__restrict__ loc, const TM *__restrict__ attn, const TV *__restrict__ grad_out,
                   GA *__restrict__ grad_value, TM *__restrict__ grad_loc, TM *__restrict__ grad_attn,
                   int S, int M, int D, int L, int Lq, int P, long long n_units)
{
    using CT = typename Compute<TV>::type;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long stride = (long long)gridDim.x * kWarps;
    const long long pix = (long long)M * D;
    for (long long qm = warp0; qm < n_units; qm += stride) {
        const int m = (int)(qm % M);
        const long long n = qm / ((long long)M * Lq);
        const long long head = n * S * pix + (long long)m * D;
        const TV *g = grad_out + qm * D;
        long long s = qm * L * P;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const long long lvl = head + start[l] * pix;
            for (int p = 0; p < P; ++p, ++s) {
                CT sx = 0, sy = 0, sa = 0;
                Taps<CT> t;
                if (make_taps<CT>(to_compute(loc[2 * s]), to_compute(loc[2 * s + 1]), H, W, t)) {
                    const CT a = to_compute(attn[s]);
                    const long long r0 = lvl + ((long long)t.y0 * W + t.x0) * pix;
                    const long long r1 = r0 + (long long)W * pix;
                    const CT w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
                    for (int c = lane; c < D; c += 32) {
                        const CT go = to_compute(g[c]);
                        const CT ga = go * a;
                        CT v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                        if (t.top && t.lef) { v1 = to_compute(value[r0 + c]);       atomic_add(grad_value + r0 + c, (GA)(w1 * ga)); }
                        if (t.top && t.rig) { v2 = to_compute(value[r0 + pix + c]); atomic_add(grad_value + r0 + pix + c, (GA)(w2 * ga)); }
                        if (t.bot && t.lef) { v3 = to_compute(value[r1 + c]);       atomic_add(grad_value + r1 + c, (GA)(w3 * ga)); }
                        if (t.bot && t.rig) { v4 = to_compute(value[r1 + pix + c]); atomic_add(grad_value + r1 + pix + c, (GA)(w4 * ga)); }
                        const CT dh = -t.hw * v1 - t.lw * v2 + t.hw * v3 + t.lw * v4;
                        const CT dw = -t.hh * v1 + t.hh * v2 - t.lh * v3 + t.lh * v4;
                        sx += (CT)W * dw * ga;
                        sy += (CT)H * dh * ga;
                        sa += go * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                    }
                }
                sx = warp_sum(sx);
                sy = warp_sum(sy);
                sa = warp_sum(sa);
                if (lane == 0) {
                    grad_loc[2 * s] = from_compute<TM, CT>(sx);
                    grad_loc[2 * s + 1] = from_compute<TM, CT>(sy);
                    grad_attn[s] = from_compute<TM, CT>(sa);
                }
            }
        }
    }
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}

int grid_for(long long warps_needed, int sm_count)
{
    long long blocks = (warps_needed + kWarps - 1) / kWarps;
    const long long cap = (long long)sm_count * 64;  // grid-stride beyond this
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename TV, typename TM>
cudaError_t fwd_launch(const Problem &p, void *out, cudaStream_t st)
{
    const int chunks = (p.D + 31) / 32;
    const long long units = (long long)p.N * p.Lq * p.M * chunks;
    if (units == 0) return cudaSuccess;
    fwd_generic_kernel<TV, TM><<<grid_for(units, p.sm_count), kThreads, 0, st>>>(
        (const TV *)p.value, p.shapes, p.start, (const TM *)p.loc, (const TM *)p.attn, (TV *)out,
        p.S, p.M, p.D, p.L, p.Lq, p.P, chunks, units);
    count_launch();
    return cudaGetLastError();
}

template <typename TV, typename TM, typename GA>
cudaError_t bwd_launch(const Problem &p, const BackwardIO &io, GA *gv_acc, cudaStream_t st)
{
    const long long n_val = (long long)p.N * p.S * p.M * p.D;
    cudaError_t e = cudaSuccess;
    if (n_val > 0) {
        e = cudaMemsetAsync(gv_acc, 0, sizeof(GA) * (size_t)n_val, st);
        if (e != cudaSuccess) return e;
    }
    const long long units = (long long)p.N * p.Lq * p.M;
    if (units == 0) return cudaSuccess;
    bwd_generic_kernel<TV, TM, GA><<<grid_for(units, p.sm_count), kThreads, 0, st>>>(
        (const TV *)p.value, p.shapes, p.start, (const TM *)p.loc, (const TM *)p.attn, (const TV *)io.grad_out, gv_acc,
        (TM *)io.grad_loc, (TM *)io.grad_attn, p.S, p.M, p.D, p.L, p.Lq, p.P, units);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t convert_f32_to_bf16(const float *src, void *dst, long long n, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    f32_to_bf16_kernel<<<sm_count * 8, 256, 0, st>>>(src, (__nv_bfloat16 *)dst, n);
    count_launch();
    return cudaGetLastError();
}

size_t bwd_generic_workspace(int dtype, int N, int S, int M, int D)
{
    return dtype == 1 /*bf16*/ ? sizeof(float) * (size_t)N * S * M * D : 0;
}

cudaError_t fwd_generic(const Problem &p, void *out, cudaStream_t st)
{
    using bf16 = __nv_bfloat16;
    if (p.dtype == 0) return fwd_launch<float, float>(p, out, st);
    if (p.dtype == 2) return fwd_launch<double, double>(p, out, st);
    if (p.loc_dtype == 1) return fwd_launch<bf16, bf16>(p, out, st);
    return fwd_launch<bf16, float>(p, out, st);
}

cudaError_t bwd_generic(const Problem &p, const BackwardIO &io, cudaStream_t st)
{
    using bf16 = __nv_bfloat16;
    if (p.dtype == 0) return bwd_launch<float, float, float>(p, io, (float *)io.grad_value, st);
    if (p.dtype == 2) return bwd_launch<double, double, double>(p, io, (double *)io.grad_value, st);
    float *acc = (float *)io.workspace;
    cudaError_t e = p.loc_dtype == 1 ? bwd_launch<bf16, bf16, float>(p, io, acc, st) : bwd_launch<bf16, float, float>(p, io, acc, st);
    if (e != cudaSuccess) return e;
    const long long n_val = (long long)p.N * p.S * p.M * p.D;
    if (n_val > 0) {
        f32_to_bf16_kernel<<<p.sm_count * 8, 256, 0, st>>>(acc, (bf16 *)io.grad_value, n_val);
        count_launch();
    }
    return cudaGetLastError();
}


namespace {
// one thread per (frame, query, head) row: grad_logit_j = a_j * (g_j - sum_k a_k g_k)   (softmax backward, in place on g)
__global__ void softmax_backward_kernel(const float *__restrict__ attn, float *__restrict__ grad, long long rows, int LP)
{
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const float *a = attn + r * LP;
        float *g = grad + r * LP;
        float dot = 0.f;
        if ((LP & 3) == 0) {
            for (int i = 0; i < LP; i += 4) {
                const float4 av = *reinterpret_cast<const float4 *>(a + i), gv = *reinterpret_cast<const float4 *>(g + i);
                dot += av.x * gv.x + av.y * gv.y + av.z * gv.z + av.w * gv.w;
            }
            for (int i = 0; i < LP; i += 4) {
                const float4 av = *reinterpret_cast<const float4 *>(a + i);
                float4 gv = *reinterpret_cast<const float4 *>(g + i);
                gv.x = av.x * (gv.x - dot); gv.y = av.y * (gv.y - dot); gv.z = av.z * (gv.z - dot); gv.w = av.w * (gv.w - dot);
                *reinterpret_cast<float4 *>(g + i) = gv;
            }
        } else {
            for (int i = 0; i < LP; ++i) dot += a[i] * g[i];
            for (int i = 0; i < LP; ++i) g[i] = a[i] * (g[i] - dot);
        }
    }
}
}  // namespace

namespace {
__global__ void zero_masked_rows_kernel(float4 *__restrict__ grad, const unsigned char *__restrict__ mask, long long rows, int row_vec4)
{
    // a warp looks at 32 x 16 mask bytes per step (one 16-byte load per lane); only rows whose byte is set are written,
    // by the whole warp -- for the decoder's all-False mask the kernel reads N*S bytes and writes nothing
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp * 512; base < rows; base += nwarps * 512) {
        const long long r0 = base + lane * 16;
        uint4 m = make_uint4(0, 0, 0, 0);
        if (r0 + 16 <= rows && ((reinterpret_cast<uintptr_t>(mask) + r0) & 15) == 0) {
            m = *reinterpret_cast<const uint4 *>(mask + r0);
        } else {
            unsigned char *mb = reinterpret_cast<unsigned char *>(&m);
            for (int i = 0; i < 16; ++i)
                if (r0 + i < rows) mb[i] = mask[r0 + i];
        }
        const unsigned any = __ballot_sync(0xffffffffu, (m.x | m.y | m.z | m.w) != 0);
        for (unsigned todo = any; todo; todo &= todo - 1) {
            const int src = __ffs(todo) - 1;
            const unsigned w[4] = {__shfl_sync(0xffffffffu, m.x, src), __shfl_sync(0xffffffffu, m.y, src),
                                   __shfl_sync(0xffffffffu, m.z, src), __shfl_sync(0xffffffffu, m.w, src)};
            for (int i = 0; i < 16; ++i) {
                if ((w[i >> 2] >> ((i & 3) * 8)) & 0xff) {
                    const long long r = base + src * 16 + i;
                    for (int c = lane; c < row_vec4; c += 32) grad[r * row_vec4 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
}
}  // namespace

cudaError_t zero_masked_rows(float *grad, const unsigned char *mask, long long rows, int row_elems, int sm_count, cudaStream_t st)
{
    if (rows <= 0 || !mask) return cudaSuccess;
    if (row_elems % 4 != 0) return cudaErrorInvalidValue;
    long long warps = (rows + 511) / 512;
    long long blocks = (warps + 7) / 8;
    if (blocks > sm_count * 8LL) blocks = sm_count * 8LL;
    zero_masked_rows_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<float4 *>(grad), mask, rows, row_elems / 4);
    count_launch();
    return cudaGetLastError();
}

cudaError_t softmax_backward_rows(const float *attn, float *grad, long long rows, int LP, int sm_count, cudaStream_t st)
{
    if (rows <= 0) return cudaSuccess;
    const int threads = 256;
    long long blocks = (rows + threads - 1) / threads;
    const long long cap = (long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    softmax_backward_kernel<<<(int)blocks, threads, 0, st>>>(attn, grad, rows, LP);
    count_launch();
    return cudaGetLastError();
}

}  // namespace msda
