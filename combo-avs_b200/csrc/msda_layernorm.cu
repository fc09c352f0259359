// msda_layernorm.cu -- row f2 of the scope table, the elementwise epilogue of the deformable encoder layer's GEMMs:
//     src = norm1(src + self_attn(...));  src = norm2(src + linear2(relu(linear1(src))))
// (reference models/modeling/pixel_decoder/msdeformattn.py:119-134), i.e. twice per layer y = LayerNorm(a + b) over rows of 256 floats,
// 41 160 rows for the S4 shape.  PyTorch runs that as an add kernel + native_layer_norm forward, and native_layer_norm_backward whose
// gamma / beta reduction alone takes 308 us per call on a B200 (measured, profiles/r08i_decoder_step_torchprof.txt: 18 % of the pixel
// decoder's training step); together with the bias-gradient column sums of the Linear layers (another 9 %) that is more time than the
// six MSDeformAttn cores.  These are HBM-bound passes: the kernels below move each tensor once.
//
//   add_layernorm forward : y = ((a + b) - mean) * rstd * gamma + beta, mean / rstd per row saved for the backward          (3 tensors moved)
//   layernorm backward    : dx = rstd * (g*gamma - mean_c(g*gamma) - xhat * mean_c(g*gamma*xhat)), xhat recomputed from a + b;
//                           dgamma = sum_rows g*xhat, dbeta = sum_rows g: per-lane partial sums over the rows a warp owns, combined per
//                           CTA in shared memory, written as one partial row per CTA and summed in a fixed order by a second small
//                           kernel -- deterministic, no atomics                                                             (4 tensors moved)
//   column sum            : out[c] = sum_rows g[r][c] (grad_bias of a Linear layer), same two-stage scheme
//
// One warp per row, lane l owns the float4 at columns r*128 + 4*l (r = 0 .. C/128 - 1): every load / store instruction of a warp covers 512
// contiguous bytes.  Persistent grid (a multiple of the SM count), rows dealt round-robin to the warps.
#include "../../include/msda_b200.h"
#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kWarps = 8;  // warps per CTA

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int R>
__global__ void __launch_bounds__(kWarps * 32)
add_layernorm_fwd_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ gamma, const float *__restrict__ beta,
                         float *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd, int M, float eps)
{
    constexpr int C = R * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 gm[R], bt[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        gm[r] = ldg4(gamma + r * 128 + lane * 4);
        bt[r] = ldg4(beta + r * 128 + lane * 4);
    }
    for (long long row = (long long)blockIdx.x * kWarps + warp; row < M; row += (long long)gridDim.x * kWarps) {
        const float *pa = a + row * C + lane * 4;
        float4 x[R];
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            x[r] = ldg4(pa + r * 128);
            if (b != nullptr) {
                const float4 t = ldg4(b + row * C + lane * 4 + r * 128);
                x[r].x += t.x; x[r].y += t.y; x[r].z += t.z; x[r].w += t.w;
            }
            s += (x[r].x + x[r].y) + (x[r].z + x[r].w);
        }
        const float mu = warp_sum(s) * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            x[r].x -= mu; x[r].y -= mu; x[r].z -= mu; x[r].w -= mu;
            q += (x[r].x * x[r].x + x[r].y * x[r].y) + (x[r].z * x[r].z + x[r].w * x[r].w);
        }
        const float rs = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float4 o;
            o.x = __fmaf_rn(x[r].x * rs, gm[r].x, bt[r].x);
            o.y = __fmaf_rn(x[r].y * rs, gm[r].y, bt[r].y);
            o.z = __fmaf_rn(x[r].z * rs, gm[r].z, bt[r].z);
            o.w = __fmaf_rn(x[r].w * rs, gm[r].w, bt[r].w);
            *reinterpret_cast<float4 *>(y + row * C + lane * 4 + r * 128) = o;
        }
        if (lane == 0) {
            mean[row] = mu;
            rstd[row] = rs;
        }
    }
}

// part[blockIdx.x][0..C) = this CTA's sum of g*xhat, part[blockIdx.x][C..2C) = its sum of g
template <int R>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_bwd_kernel(const float *__restrict__ g, const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ gamma,
                     const float *__restrict__ mean, const float *__restrict__ rstd, float *__restrict__ dx, float *__restrict__ part, int M)
{
    constexpr int C = R * 128;
    __shared__ float4 red[kWarps][2 * R][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 gm[R], dg[R], db[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        gm[r] = ldg4(gamma + r * 128 + lane * 4);
        dg[r] = db[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long row = (long long)blockIdx.x * kWarps + warp; row < M; row += (long long)gridDim.x * kWarps) {
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        float4 xh[R], gg[R];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long off = row * C + lane * 4 + r * 128;
            float4 x = ldg4(a + off);
            if (b != nullptr) {
                const float4 t = ldg4(b + off);
                x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
            }
            const float4 go = ldg4(g + off);
            xh[r] = make_float4((x.x - mu) * rs, (x.y - mu) * rs, (x.z - mu) * rs, (x.w - mu) * rs);
            gg[r] = make_float4(go.x * gm[r].x, go.y * gm[r].y, go.z * gm[r].z, go.w * gm[r].w);
            s1 += (gg[r].x + gg[r].y) + (gg[r].z + gg[r].w);
            s2 += (gg[r].x * xh[r].x + gg[r].y * xh[r].y) + (gg[r].z * xh[r].z + gg[r].w * xh[r].w);
            dg[r].x = __fmaf_rn(go.x, xh[r].x, dg[r].x); dg[r].y = __fmaf_rn(go.y, xh[r].y, dg[r].y);
            dg[r].z = __fmaf_rn(go.z, xh[r].z, dg[r].z); dg[r].w = __fmaf_rn(go.w, xh[r].w, dg[r].w);
            db[r].x += go.x; db[r].y += go.y; db[r].z += go.z; db[r].w += go.w;
        }
        const float m1 = warp_sum(s1) * (1.0f / C), m2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float4 o;
            o.x = rs * (gg[r].x - m1 - xh[r].x * m2);
            o.y = rs * (gg[r].y - m1 - xh[r].y * m2);
            o.z = rs * (gg[r].z - m1 - xh[r].z * m2);
            o.w = rs * (gg[r].w - m1 - xh[r].w * m2);
            *reinterpret_cast<float4 *>(dx + row * C + lane * 4 + r * 128) = o;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        red[warp][r][lane] = dg[r];
        red[warp][R + r][lane] = db[r];
    }
    __syncthreads();
    // 2R*32 float4 columns, summed over the warps in a fixed order
    for (int i = threadIdx.x; i < 2 * R * 32; i += kWarps * 32) {
        const int r = i >> 5, l = i & 31;
        float4 s = red[0][r][l];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
            const float4 t = red[w][r][l];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        // column of (r, l): (r % R) * 128 + 4 l, in the g*xhat half (r < R) or the g half
        float *dst = part + (size_t)blockIdx.x * 2 * C + (r < R ? 0 : C) + (r % R) * 128 + l * 4;
        *reinterpret_cast<float4 *>(dst) = s;
    }
}

// out[c] = sum over `parts` rows of part[p][c] in a fixed order; `cols` columns.  32 columns x 8 part-lanes per CTA: part-lane t adds rows
// t, t+8, ... (a serial chain of parts/8 loads instead of parts), the eight results are then added in order t = 0..7.
__global__ void __launch_bounds__(256)
sum_partials_kernel(const float *__restrict__ part, float *__restrict__ out0, float *__restrict__ out1, int parts, int cols0, int cols)
{
    __shared__ float red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (c < cols)
        for (int p = ty; p < parts; p += 8) s += part[(size_t)p * cols + c];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
#pragma unroll
        for (int t = 1; t < 8; ++t) s += red[t][tx];
        if (c < cols0) out0[c] = s;
        else if (out1 != nullptr) out1[c - cols0] = s;
    }
}

// part[blockIdx.x][0..N) = this CTA's column sums of its rows of g[M][N]; N % 4 == 0, N <= 4096
__global__ void __launch_bounds__(256)
colsum_kernel(const float *__restrict__ g, float *__restrict__ part, int M, int N)
{
    extern __shared__ float4 sred[];  // [row lanes][N/4]
    const int c4 = N >> 2;
    const int lanes = 256 / c4 > 0 ? 256 / c4 : 1;  // threads that share one float4 column, each on its own rows
    const int col = threadIdx.x % c4, rl = threadIdx.x / c4;
    // N/4 > 256: every thread walks several float4 columns
    for (int cb = 0; cb < c4; cb += 256) {
        const int cc = c4 > 256 ? cb + threadIdx.x : col;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cc < c4 && (c4 > 256 || rl < lanes)) {
            const int step = (c4 > 256 ? 1 : lanes) * gridDim.x, first = blockIdx.x * (c4 > 256 ? 1 : lanes) + (c4 > 256 ? 0 : rl);
            for (long long row = first; row < M; row += step) {
                const float4 t = ldg4(g + row * N + cc * 4);
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
        }
        if (c4 > 256) {
            if (cc < c4) *reinterpret_cast<float4 *>(part + (size_t)blockIdx.x * N + cc * 4) = s;
        } else {
            if (rl < lanes) sred[rl * c4 + col] = s;
            __syncthreads();
            if (threadIdx.x < c4) {
                float4 acc = sred[threadIdx.x];
                for (int k = 1; k < lanes; ++k) {
                    const float4 t = sred[k * c4 + threadIdx.x];
                    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                }
                *reinterpret_cast<float4 *>(part + (size_t)blockIdx.x * N + threadIdx.x * 4) = acc;
            }
            break;
        }
    }
}

int ln_grid(int M, int sm_count)
{
    long long want = ((long long)M + kWarps - 1) / kWarps;
    const long long cap = (long long)sm_count * 4;  // 4 CTAs of 8 warps per SM: enough loads in flight for an HBM-bound pass
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace

bool layernorm_supported(int C) { return C >= 128 && C <= 512 && C % 128 == 0; }  // (the backward's static shared memory: 8 KB per 128 columns)

size_t layernorm_backward_workspace(int M, int C, int sm_count) { return (size_t)ln_grid(M, sm_count) * 2 * C * sizeof(float); }

cudaError_t add_layernorm_forward(const float *a, const float *b, const float *gamma, const float *beta, float *y, float *mean, float *rstd,
                                  int M, int C, float eps, int sm_count, cudaStream_t st)
{
    const int grid = ln_grid(M, sm_count);
    switch (C / 128) {
#define MSDA_LN_CASE(R) case R: add_layernorm_fwd_kernel<R><<<grid, kWarps * 32, 0, st>>>(a, b, gamma, beta, y, mean, rstd, M, eps); break;
        MSDA_LN_CASE(1) MSDA_LN_CASE(2) MSDA_LN_CASE(3) MSDA_LN_CASE(4)
#undef MSDA_LN_CASE
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t layernorm_backward(const float *g, const float *a, const float *b, const float *gamma, const float *mean, const float *rstd, float *dx,
                               float *dgamma, float *dbeta, float *workspace, int M, int C, int sm_count, cudaStream_t st)
{
    const int grid = ln_grid(M, sm_count);
    switch (C / 128) {
#define MSDA_LN_CASE(R) case R: layernorm_bwd_kernel<R><<<grid, kWarps * 32, 0, st>>>(g, a, b, gamma, mean, rstd, dx, workspace, M); break;
        MSDA_LN_CASE(1) MSDA_LN_CASE(2) MSDA_LN_CASE(3) MSDA_LN_CASE(4)
#undef MSDA_LN_CASE
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_partials_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(workspace, dgamma, dbeta, grid, C, 2 * C);
    count_launch(2);
    return cudaGetLastError();
}

int colsum_grid(int M, int N, int sm_count)
{
    const int c4 = N >> 2, lanes = c4 > 256 ? 1 : 256 / c4;
    long long want = ((long long)M + lanes - 1) / lanes;
    const long long cap = (long long)sm_count * 4;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

bool colsum_supported(int N) { return N >= 4 && N % 4 == 0 && N <= 4096; }

size_t colsum_workspace(int M, int N, int sm_count) { return (size_t)colsum_grid(M, N, sm_count) * N * sizeof(float); }

cudaError_t colsum(const float *g, float *out, float *workspace, int M, int N, int sm_count, cudaStream_t st)
{
    const int grid = colsum_grid(M, N, sm_count), c4 = N >> 2, lanes = c4 > 256 ? 1 : 256 / c4;
    colsum_kernel<<<grid, 256, c4 > 256 ? 0 : (size_t)lanes * c4 * sizeof(float4), st>>>(g, workspace, M, N);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sum_partials_kernel<<<(N + 31) / 32, 256, 0, st>>>(workspace, out, nullptr, grid, N, N);
    count_launch(2);
    return cudaGetLastError();
}

}  // namespace msda
