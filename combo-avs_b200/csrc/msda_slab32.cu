// msda_slab32.cu -- D == 32, float32, geometry known on the host: shared-memory SLAB kernels.
//
// Why: the op gathers 4*L*P = 48 rows of 128 B to produce one 128 B output row.  Served from L2 that is 2.02 GB per
// S4 forward and the L2 fabric is the limit; one (frame, head) slab is only S*128 B = 131.7 KB, so a CTA stages the
// slab once in shared memory (zero-padded by one pixel all around each level, so all four neighbours of any
// in-range sample exist and the corner predicates of ms_deform_im2col_cuda.cuh:58-82 disappear) and gathers at the
// 128 B/clk/SM shared-memory rate instead.
//
// Mapping: SPLIT=2 lanes per (query, head): each lane owns 16 of the 32 channels (four 16-byte chunks) and walks
// all L*P samples itself, so there is no cross-lane reduction at all.  Lane (q, half) reads chunk
// half*4 + ((k + q) & 3) in its k-th load: any 8 consecutive lanes (one LDS.128 phase) hit 8 different chunks =
// all 32 banks exactly once, whatever rows they read -> conflict-free random row gather.
//
// Scheduling: work unit = (slab, pass) where a pass is a contiguous block of <= QPP queries; the host picks the
// number of passes per slab that balances N*M*passes units over the SMs; each persistent CTA (one per SM) owns a
// contiguous unit range, so consecutive units reuse the resident slab.
#include <cmath>

#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kD = 32;
constexpr int kRowBytes = kD * 4;

struct SlabParams {
    int L, P, S, M, Lq;
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    int prow0[kMaxLevels];    // first padded row of the level inside the slab; padded pitch is W+2
    float hmax[kMaxLevels];   // largest float below H (clamp bound that leaves in-range samples untouched)
    float wmax[kMaxLevels];
    int padded_rows;          // total rows of the padded slab
    int qpp;                  // queries per pass
    int pps;                  // passes per slab
    long long units;          // N*M*pps
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// copy value[n, :, m, :] into the padded slab (interior only; the border was zeroed once)
__device__ __forceinline__ void load_slab(float4 *slab, const float *__restrict__ value, const SlabParams &sp, long long nm)
{
    const long long n = nm / sp.M;
    const int m = (int)(nm - n * sp.M);
    const float *src = value + ((n * sp.S) * sp.M + m) * kD;
    const int pixf = sp.M * kD;
    for (int l = 0; l < sp.L; ++l) {
        const int W = sp.W[l], cnt = sp.H[l] * W * 8;
        const float *lsrc = src + (long long)sp.start[l] * pixf;
        float4 *ldst = slab + (size_t)(sp.prow0[l] + (W + 2) + 1) * 8;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int pix = i >> 3, c = i & 7;
            const int y = pix / W, x = pix - y * W;
            cp_async16(ldst + (size_t)(y * (W + 2) + x) * 8 + c, lsrc + (long long)pix * pixf + c * 4);
        }
    }
    cp_async_wait_all();
}

template <int P_T>
__global__ void __launch_bounds__(512, 1)
fwd_slab32_kernel(const float *__restrict__ value, const float *__restrict__ loc, const float *__restrict__ attn,
                  float *__restrict__ out, const __grid_constant__ SlabParams sp)
{
    extern __shared__ float4 slab[];
    const int P = P_T > 0 ? P_T : sp.P;
    for (int i = threadIdx.x; i < sp.padded_rows * 8; i += blockDim.x) slab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long u_lo = sp.units * blockIdx.x / gridDim.x, u_hi = sp.units * (blockIdx.x + 1) / gridDim.x;
    const int pair = threadIdx.x >> 1, half = threadIdx.x & 1;
    unsigned off[4];  // byte offset of this lane's k-th chunk inside a row
#pragma unroll
    for (int k = 0; k < 4; ++k) off[k] = (half * 4 + ((k + pair) & 3)) * 16;
    const unsigned slab_s = (unsigned)__cvta_generic_to_shared(slab);
    const int LP = sp.L * P;
    long long resident = -1;
    for (long long u = u_lo; u < u_hi; ++u) {
        const long long nm = u / sp.pps;
        const int pass = (int)(u - nm * sp.pps);
        if (nm != resident) {
            __syncthreads();  // everyone finished with the previous slab (and with the zero fill)
            load_slab(slab, value, sp, nm);
            __syncthreads();
            resident = nm;
        }
        const int q = pass * sp.qpp + pair;
        if (pair >= sp.qpp || q >= sp.Lq) continue;
        const long long n = nm / sp.M;
        const int m = (int)(nm - n * sp.M);
        const long long qm = (n * sp.Lq + q) * sp.M + m;
        const float2 *loc2 = reinterpret_cast<const float2 *>(loc) + qm * LP;
        const float *att = attn + qm * LP;
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < sp.L; ++l) {
            const float Hf = (float)sp.H[l], Wf = (float)sp.W[l];
            const float hmax = sp.hmax[l], wmax = sp.wmax[l];
            const int pitch = sp.W[l] + 2;
            const unsigned lvl_s = slab_s + (unsigned)(sp.prow0[l] + pitch + 1) * kRowBytes;  // address of pixel (0,0)
            const unsigned pitch_b = (unsigned)pitch * kRowBytes;
#pragma unroll
            for (int p = 0; p < (P_T > 0 ? P_T : 1); ++p) {
                for (int pp = (P_T > 0 ? p : 0); pp < (P_T > 0 ? p + 1 : P); ++pp) {
                    const int s = l * P + pp;
                    const float2 xy = __ldg(loc2 + s);
                    float a = __ldg(att + s);
                    const float h = __fmaf_rn(xy.y, Hf, -0.5f), w = __fmaf_rn(xy.x, Wf, -0.5f);
                    const bool ok = h > -1.f && w > -1.f && h < Hf && w < Wf;
                    a = ok ? a : 0.f;
                    const float hc = fminf(fmaxf(h, -1.f), hmax), wc = fminf(fmaxf(w, -1.f), wmax);
                    const float fh = floorf(hc), fw = floorf(wc);
                    const float lh = hc - fh, lw = wc - fw;
                    const float ahh = a * (1.f - lh), alh = a * lh, hw = 1.f - lw;
                    const float w1 = ahh * hw, w2 = ahh * lw, w3 = alh * hw, w4 = alh * lw;
                    const unsigned r00 = lvl_s + (unsigned)((int)fh * pitch + (int)fw) * kRowBytes;
                    const unsigned r10 = r00 + pitch_b;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float4 v1, v2, v3, v4;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(r00 + off[k]));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+128];" : "=f"(v2.x), "=f"(v2.y), "=f"(v2.z), "=f"(v2.w) : "r"(r00 + off[k]));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v3.x), "=f"(v3.y), "=f"(v3.z), "=f"(v3.w) : "r"(r10 + off[k]));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+128];" : "=f"(v4.x), "=f"(v4.y), "=f"(v4.z), "=f"(v4.w) : "r"(r10 + off[k]));
                        acc[k].x = fmaf(w1, v1.x, fmaf(w2, v2.x, fmaf(w3, v3.x, fmaf(w4, v4.x, acc[k].x))));
                        acc[k].y = fmaf(w1, v1.y, fmaf(w2, v2.y, fmaf(w3, v3.y, fmaf(w4, v4.y, acc[k].y))));
                        acc[k].z = fmaf(w1, v1.z, fmaf(w2, v2.z, fmaf(w3, v3.z, fmaf(w4, v4.z, acc[k].z))));
                        acc[k].w = fmaf(w1, v1.w, fmaf(w2, v2.w, fmaf(w3, v3.w, fmaf(w4, v4.w, acc[k].w))));
                    }
                }
            }
        }
        float *o = out + qm * kD;
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<float4 *>(o + (off[k] >> 2)) = acc[k];
    }
}

// choose passes-per-slab so that N*M*pps units spread evenly over `ctas` persistent CTAs
void plan_passes(int Lq, long long slabs, int ctas, int max_pairs, int &qpp, int &pps)
{
    const int min_pps = (Lq + max_pairs - 1) / max_pairs;
    double best = 1e300;
    qpp = (Lq + min_pps - 1) / min_pps;
    pps = min_pps;
    for (int c = min_pps; c <= min_pps + 24; ++c) {
        const int q = (Lq + c - 1) / c;
        const long long rounds = (slabs * c + ctas - 1) / ctas;
        // cost ~ rounds * (pass time + a per-pass overhead); prefer fewer passes on ties
        const double cost = (double)rounds * (q + 6.0);
        if (cost < best - 1e-9) {
            best = cost;
            qpp = q;
            pps = c;
        }
    }
}

bool make_params(const Problem &p, SlabParams &sp, size_t &smem_bytes, int &threads, int &grid)
{
    if (!(p.dtype == 0 && p.loc_dtype == 0 && p.D == kD && p.have_host && p.L <= kMaxLevels)) return false;
    sp.L = p.L; sp.P = p.P; sp.S = p.S; sp.M = p.M; sp.Lq = p.Lq;
    int rows = 0;
    for (int l = 0; l < p.L; ++l) {
        sp.H[l] = p.table.H[l];
        sp.W[l] = p.table.W[l];
        sp.start[l] = p.table.start[l];
        sp.prow0[l] = rows;
        sp.hmax[l] = nextafterf((float)sp.H[l], 0.f);
        sp.wmax[l] = nextafterf((float)sp.W[l], 0.f);
        rows += (sp.H[l] + 2) * (sp.W[l] + 2);
        if (rows > (1 << 20)) return false;
    }
    sp.padded_rows = rows;
    smem_bytes = (size_t)rows * kRowBytes;
    if (smem_bytes > (size_t)p.smem_optin) return false;
    const long long slabs = (long long)p.N * p.M;
    grid = (int)(slabs * 16 < p.sm_count ? slabs * 16 : p.sm_count);
    if (grid < 1) grid = 1;
    plan_passes(p.Lq, slabs, grid, 256, sp.qpp, sp.pps);
    sp.units = slabs * sp.pps;
    threads = ((sp.qpp * 2 + 31) / 32) * 32;
    if (threads > 512) threads = 512;
    return true;
}

}  // namespace

bool slab32_supported(const Problem &p)
{
    SlabParams sp;
    size_t smem;
    int threads, grid;
    return make_params(p, sp, smem, threads, grid);
}

cudaError_t fwd_slab32(const Problem &p, void *out, cudaStream_t st)
{
    SlabParams sp;
    size_t smem;
    int threads, grid;
    if (!make_params(p, sp, smem, threads, grid)) return cudaErrorInvalidValue;
    if ((long long)p.N * p.M * p.Lq == 0) return cudaSuccess;
    auto kern = p.P == 4 ? fwd_slab32_kernel<4> : fwd_slab32_kernel<0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, smem, st>>>((const float *)p.value, (const float *)p.loc, (const float *)p.attn, (float *)out, sp);
    count_launch();
    return cudaGetLastError();
}

}  // namespace msda
