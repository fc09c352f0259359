// msda_warp32.cu -- D == 32, float32: warp-cooperative kernels that gather straight from L1/L2.
//
// Mapping ("variant A" of DESIGN.md): one warp per (frame n, head m, query q).  A 32-channel fp32 row of the
// channel-contiguous value tensor is exactly one 128-byte line, so 8 lanes x float4 cover a row with one LDG.128
// each and the 4 lane-groups of the warp take 4 different samples of the (query, head) at the same time: every
// warp-level load instruction moves four full 128-byte lines.  Partial sums are combined with warp shuffles.
//
// Scheduling: persistent CTAs, one per SM, each owning a contiguous range of units in (n, m, q) order -- i.e.
// HEAD-MAJOR, not the reference's channel-fastest flat index (ms_deform_im2col_cuda.cuh:258-268).  All warps of
// an SM then gather from one (frame, head) slab at a time (S*128 B = 131.7 KB for the S4 pyramid), which fits the
// L1 when the shared-memory carve-out is set to zero, instead of from a whole 1 MB frame.
//
// Geometry (H_l, W_l, start_l) is read from DEVICE memory (the int64 tensors PyTorch hands over) and staged in
// shared memory once per CTA, so these kernels need no host mirror and are the default "no sync" path.
//
// Backward: grad_value through 16-byte vector reductions (red.global.add.v4.f32 -> REDG.E.ADD.F32x4): 4x fewer
// L2 atomic operations than the reference's scalar atomicAdd (ms_deform_im2col_cuda.cuh:130-157);
// grad_sampling_loc / grad_attn_weight are 8-lane shuffle reductions with a single plain store, replacing the
// reference's shared-memory cache + serial thread-0 sum + 24 __syncthreads per block (:352-405).
#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kD = 32;
constexpr int kFwdThreads = 1024;
constexpr int kBwdThreads = 512;

struct Geo {
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
};

__device__ __forceinline__ void stage_geometry(Geo &g, const int64_t *shapes, const int64_t *start, int L)
{
    if (threadIdx.x < L) {
        g.H[threadIdx.x] = (int)shapes[2 * threadIdx.x];
        g.W[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
        g.start[threadIdx.x] = (int)start[threadIdx.x];
    }
    __syncthreads();
}

// unit u in [0, N*M*Lq) is (n, m, q) with q fastest; CTA b owns [b*U/G, (b+1)*U/G)
__device__ __forceinline__ void cta_range(long long U, long long &lo, long long &hi)
{
    lo = U * blockIdx.x / gridDim.x;
    hi = U * (blockIdx.x + 1) / gridDim.x;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fwd_warp32_kernel(const float *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
                  const float *__restrict__ loc, const float *__restrict__ attn, float *__restrict__ out,
                  int S, int M, int L, int Lq, int P, long long U)
{
    __shared__ Geo geo;
    stage_geometry(geo, shapes, start, L);
    const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7;
    const int LP = L * P;
    const int pix = M * kD;  // floats between neighbouring pixels
    long long lo, hi;
    cta_range(U, lo, hi);
    for (long long u = lo + (threadIdx.x >> 5); u < hi; u += THREADS / 32) {
        const int q = (int)(u % Lq);
        const long long nm = u / Lq;
        const int m = (int)(nm % M);
        const long long n = nm / M;
        const long long qm = (n * Lq + q) * M + m;
        const float *vhead = value + (n * S * M + m) * kD + j * 4;
        const float2 *loc2 = reinterpret_cast<const float2 *>(loc) + qm * LP;
        const float *att = attn + qm * LP;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = grp; s < LP; s += 4) {
            const int l = s / P;
            const int H = geo.H[l], W = geo.W[l];
            const float2 xy = __ldg(loc2 + s);
            const float a = __ldg(att + s);
            Taps<float> t;
            if (!make_taps<float>(xy.x, xy.y, H, W, t)) continue;
            const float *r0 = vhead + (long long)(geo.start[l] + t.y0 * W + t.x0) * pix;
            const float *r1 = r0 + (long long)W * pix;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 v1 = (t.top && t.lef) ? ldg4(r0) : z;
            const float4 v2 = (t.top && t.rig) ? ldg4(r0 + pix) : z;
            const float4 v3 = (t.bot && t.lef) ? ldg4(r1) : z;
            const float4 v4 = (t.bot && t.rig) ? ldg4(r1 + pix) : z;
            const float w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
            acc.x += (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x) * a;
            acc.y += (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y) * a;
            acc.z += (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z) * a;
            acc.w += (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w) * a;
        }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if (grp == 0) *reinterpret_cast<float4 *>(out + qm * kD + j * 4) = acc;
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
bwd_warp32_kernel(const float *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
                  const float *__restrict__ loc, const float *__restrict__ attn, const float *__restrict__ grad_out,
                  float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn,
                  int S, int M, int L, int Lq, int P, long long U)
{
    __shared__ Geo geo;
    stage_geometry(geo, shapes, start, L);
    const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7;
    const int LP = L * P;
    const int LPr = (LP + 3) & ~3;  // every lane runs the same number of iterations (shuffles need the full warp)
    const int pix = M * kD;
    long long lo, hi;
    cta_range(U, lo, hi);
    for (long long u = lo + (threadIdx.x >> 5); u < hi; u += THREADS / 32) {
        const int q = (int)(u % Lq);
        const long long nm = u / Lq;
        const int m = (int)(nm % M);
        const long long n = nm / M;
        const long long qm = (n * Lq + q) * M + m;
        const long long head = (n * S * M + m) * kD + j * 4;
        const float2 *loc2 = reinterpret_cast<const float2 *>(loc) + qm * LP;
        const float *att = attn + qm * LP;
        const float4 g = ldg4(grad_out + qm * kD + j * 4);
        for (int s = grp; s < LPr; s += 4) {
            float sx = 0.f, sy = 0.f, sa = 0.f;
            const bool have = s < LP;
            if (have) {
                const int l = s / P;
                const int H = geo.H[l], W = geo.W[l];
                const float2 xy = __ldg(loc2 + s);
                const float a = __ldg(att + s);
                Taps<float> t;
                if (make_taps<float>(xy.x, xy.y, H, W, t)) {
                    const long long r0 = head + (long long)(geo.start[l] + t.y0 * W + t.x0) * pix;
                    const long long r1 = r0 + (long long)W * pix;
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    const bool c1 = t.top && t.lef, c2 = t.top && t.rig, c3 = t.bot && t.lef, c4 = t.bot && t.rig;
                    const float4 v1 = c1 ? ldg4(value + r0) : z;
                    const float4 v2 = c2 ? ldg4(value + r0 + pix) : z;
                    const float4 v3 = c3 ? ldg4(value + r1) : z;
                    const float4 v4 = c4 ? ldg4(value + r1 + pix) : z;
                    const float w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
                    const float4 ga = make_float4(g.x * a, g.y * a, g.z * a, g.w * a);
                    if (c1) red_add_v4(grad_value + r0, w1 * ga.x, w1 * ga.y, w1 * ga.z, w1 * ga.w);
                    if (c2) red_add_v4(grad_value + r0 + pix, w2 * ga.x, w2 * ga.y, w2 * ga.z, w2 * ga.w);
                    if (c3) red_add_v4(grad_value + r1, w3 * ga.x, w3 * ga.y, w3 * ga.z, w3 * ga.w);
                    if (c4) red_add_v4(grad_value + r1 + pix, w4 * ga.x, w4 * ga.y, w4 * ga.z, w4 * ga.w);
#define MSDA_CH(f)                                                                     \
    {                                                                                  \
        const float dh = -t.hw * v1.f - t.lw * v2.f + t.hw * v3.f + t.lw * v4.f;       \
        const float dw = -t.hh * v1.f + t.hh * v2.f - t.lh * v3.f + t.lh * v4.f;       \
        sx += (float)W * dw * ga.f;                                                    \
        sy += (float)H * dh * ga.f;                                                    \
        sa += g.f * (w1 * v1.f + w2 * v2.f + w3 * v3.f + w4 * v4.f);                   \
    }
                    MSDA_CH(x) MSDA_CH(y) MSDA_CH(z) MSDA_CH(w)
#undef MSDA_CH
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
                sa += __shfl_xor_sync(0xffffffffu, sa, o);
            }
            if (have && j == 0) {
                *reinterpret_cast<float2 *>(grad_loc + (qm * LP + s) * 2) = make_float2(sx, sy);
                grad_attn[qm * LP + s] = sa;
            }
        }
    }
}

}  // namespace

bool warp32_supported(const Problem &p)
{
    return p.dtype == 0 && p.loc_dtype == 0 && p.D == kD && p.L <= kMaxLevels;
}

cudaError_t fwd_warp32(const Problem &p, void *out, cudaStream_t st)
{
    const long long U = (long long)p.N * p.M * p.Lq;
    if (U == 0) return cudaSuccess;
    static const cudaError_t attr = cudaFuncSetAttribute(fwd_warp32_kernel<kFwdThreads>,
                                                         cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    (void)attr;
    const long long warps = kFwdThreads / 32;
    int grid = (int)((U + warps - 1) / warps < p.sm_count ? (U + warps - 1) / warps : p.sm_count);
    fwd_warp32_kernel<kFwdThreads><<<grid, kFwdThreads, 0, st>>>(
        (const float *)p.value, p.shapes, p.start, (const float *)p.loc, (const float *)p.attn, (float *)out,
        p.S, p.M, p.L, p.Lq, p.P, U);
    count_launch();
    return cudaGetLastError();
}

cudaError_t bwd_warp32(const Problem &p, const BackwardIO &io, cudaStream_t st)
{
    const long long n_val = (long long)p.N * p.S * p.M * p.D;
    if (n_val > 0) {
        cudaError_t e = cudaMemsetAsync(io.grad_value, 0, sizeof(float) * (size_t)n_val, st);
        if (e != cudaSuccess) return e;
    }
    const long long U = (long long)p.N * p.M * p.Lq;
    if (U == 0) return cudaSuccess;
    static const cudaError_t attr = cudaFuncSetAttribute(bwd_warp32_kernel<kBwdThreads>,
                                                         cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    (void)attr;
    const long long warps = kBwdThreads / 32;
    long long want = (U + warps - 1) / warps;
    int grid = (int)(want < 2LL * p.sm_count ? want : 2LL * p.sm_count);
    bwd_warp32_kernel<kBwdThreads><<<grid, kBwdThreads, 0, st>>>(
        (const float *)p.value, p.shapes, p.start, (const float *)p.loc, (const float *)p.attn, (const float *)io.grad_out,
        (float *)io.grad_value, (float *)io.grad_loc, (float *)io.grad_attn, p.S, p.M, p.L, p.Lq, p.P, U);
    count_launch();
    return cudaGetLastError();
}

}  // namespace msda
