// msda_scatter32.cu -- D == 32, float32, geometry known on the host: the SCATTER half of the split backward (grad_value only).
//
// Reference: ms_deform_im2col_cuda.cuh:306-408 adds w_i * attn * grad_out to the four neighbours of every sample with scalar
// atomicAdd (cuh:130-157) -- 505 M atomics for the S4 shape.  Measured on the B200 (profiles/r01_microbench_pipes.txt) any
// scatter with atomics, global or shared, costs >= 337 us, so grad_value is accumulated by SORT-AND-SEGMENT, as in
// msda_sorted32.cu: the samples of one (frame, head, level) are counting-sorted by bilinear cell in shared memory; 8-lane groups
// walk contiguous ranges of the sorted list, keep the cell's four row sums in registers in the basis (1, lw, lh, lh*lw), and
// leave with one red.global.add.v4.f32 per neighbour row per cell visit.
//
// What is different from msda_sorted32.cu (which also computes grad_sampling_loc / grad_attn_weight in the same walk):
//   * nothing here needs the value tensor -- no 115 KB level slab in shared memory, no dot products, no 8-lane reductions, no
//     per-sample result stores (msda_slab32.cu's gather kernel produces those at the shared-memory gather rate instead): the walk is
//     ~21 instructions per 4 samples instead of ~72, 56 registers instead of 80;
//   * shared memory per CTA is the sorted list only (<= 70 KB), so THREE CTAs share an SM and the latency-bound sort phases of one
//     run under the walks of the others (the single-CTA sort phases were 22 % of msda_sorted32's time);
//   * work is split over the CTAs by QUERY RANGE of the (frame, head, level) sequence, not by unit: perfectly balanced for any
//     N, no unit is sorted twice, and a CTA's piece is sorted on its own (grad_value accumulates with red.global either way).
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kD = 32;
constexpr int kMaxCells = 1024;     // cell-grid size per level

struct ScatterParams {
    int L, P, S, M, Lq;
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    int cshift[kMaxLevels];  // log2 of the cell-grid pitch (power of two >= W+1): cell = (y0+1) << cshift | (x0+1)
    float rw[kMaxLevels], rh[kMaxLevels];  // 1/W, 1/H (fused prologue)
    int qmax;                // queries per piece (<= Lq): bounds the sorted list
    long long total_q;       // N*M*L*Lq: the query sequence that is cut into gridDim.x equal ranges
    int fused;               // `loc` holds raw offsets: location = ref + offset * (1/W, 1/H)
    int probe;               // profiling switches (MSDA_SCATTER_PROBE): 1 = sort phases only, 2 = no flush
};

struct f4p {
    float2 lo, hi;
};
__device__ __forceinline__ f4p pk(float4 v) { return {make_float2(v.x, v.y), make_float2(v.z, v.w)}; }
__device__ __forceinline__ f4p psub(f4p a, f4p b)
{
    const float2 m1 = make_float2(-1.f, -1.f);
    return {__ffma2_rn(b.lo, m1, a.lo), __ffma2_rn(b.hi, m1, a.hi)};
}
__device__ __forceinline__ float floor_magic(float x, int &i)
{
    const float t = __fadd_rd(x, 12582912.f);  // valid for |x| < 2^22
    i = __float_as_int(t) - 0x4B400000;
    return t - 12582912.f;
}
__device__ __forceinline__ void pin(int &x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(unsigned &x) { asm volatile("" : "+r"(x)); }
template <typename T> __device__ __forceinline__ void pin(T *&x) { asm volatile("" : "+l"(x)); }
__device__ __forceinline__ float4 lds4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int P_T, bool FUSED, int kThreads, int kCtasPerSm>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
bwd_scatter32_kernel(const float *__restrict__ loc, const float *__restrict__ attn, const float *__restrict__ grad_out,
                     float *__restrict__ grad_value, const float *__restrict__ ref, const __grid_constant__ ScatterParams sp)
{
    constexpr int kGroups = kThreads / 8;
    constexpr int kCPT = (kMaxCells + kThreads - 1) / kThreads;  // cells scanned per thread
    extern __shared__ float4 smem4[];
    const int Lq = sp.Lq, P = P_T > 0 ? P_T : sp.P, LP = sp.L * P;
    float4 *ent = smem4;                                   // sorted samples: {lh, lw, attn, cell << 16 | query}
    int *cell_off = reinterpret_cast<int *>(ent + sp.qmax * P + 4);  // (four spare entries, see the prefetch) kMaxCells + 1 cursors
    int *warp_tot = cell_off + kMaxCells + 8;
    // per-cell facts of the piece's level: row offset of the cell's top-left neighbour (biased by W+1 so that it is never negative)
    // in the low 20 bits, 4-bit mask of the neighbours inside the map above them
    unsigned cinfo_s = (unsigned)__cvta_generic_to_shared(warp_tot + 32);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int grp = tid >> 3;
    const int chf = (tid & 7) * 4;                         // this lane's four channels of a 32-channel row
    unsigned ent_s = (unsigned)__cvta_generic_to_shared(ent);
    pin(ent_s);
    const int pix = sp.M * kD;
    const int qstride = sp.M * LP;  // samples between consecutive queries of one head

    long long cursor = sp.total_q * blockIdx.x / gridDim.x;
    const long long q_end = sp.total_q * (blockIdx.x + 1) / gridDim.x;
    while (cursor < q_end) {
        // ---- next piece: queries [q_lo, q_hi) of unit u = (n*M + m)*L + l --------------------------------------------------------
        const int u = (int)(cursor / Lq);
        const int q_lo = (int)(cursor - (long long)u * Lq);
        int q_hi = q_lo + sp.qmax;
        if (q_hi > Lq) q_hi = Lq;
        if (q_hi - q_lo > q_end - cursor) q_hi = q_lo + (int)(q_end - cursor);
        cursor += q_hi - q_lo;
        const int EP = (q_hi - q_lo) * P;
        const int nm = u / sp.L;
        const int l = u - nm * sp.L;
        const long long n = nm / sp.M;
        const int m = nm - (int)n * sp.M;
        const int H = sp.H[l], W = sp.W[l], cs = sp.cshift[l], cells = (H + 1) << cs;
        const float Hf = (float)H, Wf = (float)W;
        const long long meta0 = ((n * Lq + q_lo) * sp.M + m) * (long long)LP + l * P;  // sample (q_lo, p = 0) of this (n, m, l)
        const float *locb = loc + 2 * meta0;
        const float *att = attn + meta0;
        const float *refb = FUSED ? ref + ((n * Lq + q_lo) * sp.L + l) * 2 : nullptr;
        const float rw = sp.rw[l], rh = sp.rh[l];
        // raw location of sample i of the piece (the loads of a batch are issued together: one L2 round trip per batch, not per sample)
        auto load_xy = [&](int i) -> float2 {
            const int q = P_T > 0 ? i / P_T : i / P, p = i - q * P;
            float2 xy = __ldg(reinterpret_cast<const float2 *>(locb + 2 * (q * qstride + p)));
            if (FUSED) {  // module prologue (ms_deform_attn.py:110-112), the same FMA as the forward and the gather kernel
                const float2 rf = __ldg(reinterpret_cast<const float2 *>(refb + q * sp.L * 2));
                xy.x = __fmaf_rn(xy.x, rw, rf.x);
                xy.y = __fmaf_rn(xy.y, rh, rf.y);
            }
            return xy;
        };
        // cell of a location (or -1 when it lies outside the map: cuh:352-356) and its fractions
        auto locate = [&](float2 xy, float &lh, float &lw) -> int {
            const float h = __fmaf_rn(xy.y, Hf, -0.5f), w = __fmaf_rn(xy.x, Wf, -0.5f);
            if (!(h > -1.f && w > -1.f && h < Hf && w < Wf)) return -1;
            int y0, x0;
            lh = h - floor_magic(h, y0);
            lw = w - floor_magic(w, x0);
            return ((y0 + 1) << cs) + x0 + 1;
        };
        constexpr int kBatch = 4;
        __syncthreads();  // the previous piece is completely finished with ent / cell_off
        for (int i = tid; i <= cells; i += kThreads) cell_off[i] = 0;
        for (int c = tid; c < cells; c += kThreads) {
            const int cy = c >> cs, cx = c & ((1 << cs) - 1);
            const int inm = ((cy >= 1 && cx >= 1) ? 1 : 0) | ((cy >= 1 && cx <= W - 1) ? 2 : 0) |
                            ((cy <= H - 1 && cx >= 1) ? 4 : 0) | ((cy <= H - 1 && cx <= W - 1) ? 8 : 0);
            (warp_tot + 32)[c] = (cy * W + cx) | (inm << 20);  // ((cy-1)*W + cx-1) + (W+1)
        }
        __syncthreads();
        // ---- pass 1: histogram of the cells (shared-memory integer reductions, no return value) -----------------------------------
        for (int i0 = tid; i0 < EP; i0 += kBatch * kThreads) {
            float2 xy[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) xy[b] = i0 + b * kThreads < EP ? load_xy(i0 + b * kThreads) : make_float2(2.f, 2.f);
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                float lh, lw;
                const int cell = locate(xy[b], lh, lw);
                if (i0 + b * kThreads < EP && cell >= 0) atomicAdd(&cell_off[cell], 1);
            }
        }
        __syncthreads();
        // ---- exclusive scan of the cell counters (<= kMaxCells cells, kCPT consecutive cells per thread) ---------------------------
        {
            int v[kCPT], sum = 0;
#pragma unroll
            for (int k = 0; k < kCPT; ++k) {
                const int i = kCPT * tid + k;
                v[k] = i < cells ? cell_off[i] : 0;
                sum += v[k];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[wid] = incl;
            __syncthreads();
            if (wid == 0) {
                int t = lane < kThreads / 32 ? warp_tot[lane] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int x = __shfl_up_sync(0xffffffffu, t, o);
                    if (lane >= o) t += x;
                }
                if (lane < kThreads / 32) warp_tot[lane] = t;  // inclusive totals of warps
            }
            __syncthreads();
            int base = (wid > 0 ? warp_tot[wid - 1] : 0) + incl - sum;
#pragma unroll
            for (int k = 0; k < kCPT; ++k) {
                const int i = kCPT * tid + k;
                if (i < cells) cell_off[i] = base;
                base += v[k];
            }
            if (tid == kThreads - 1) {
                const int tot = warp_tot[kThreads / 32 - 1];
                cell_off[kMaxCells] = tot;  // number of in-range samples
                // the grad_out prefetch runs up to two entries past a group's range without clamping: past the list's end it must
                // still find a valid query index (0)
                for (int k = 0; k < 4; ++k) ent[tot + k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        // ---- pass 2: every in-range sample takes the next slot of its cell (the counters become per-cell cursors) -------------------
        for (int i0 = tid; i0 < EP; i0 += kBatch * kThreads) {
            float2 xy[kBatch];
            float av[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int i = i0 + b * kThreads;
                xy[b] = make_float2(2.f, 2.f);
                av[b] = 0.f;
                if (i < EP) {
                    const int q = P_T > 0 ? i / P_T : i / P, p = i - q * P;
                    xy[b] = load_xy(i);
                    av[b] = __ldg(att + q * qstride + p);
                }
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int i = i0 + b * kThreads;
                float lh = 0.f, lw = 0.f;
                const int cell = locate(xy[b], lh, lw);
                if (i < EP && cell >= 0) {
                    const int q = P_T > 0 ? i / P_T : i / P;
                    const int slot = atomicAdd(&cell_off[cell], 1);
                    ent[slot] = make_float4(lh, lw, av[b], __int_as_float((cell << 16) | q));
                }
            }
        }
        __syncthreads();  // sorted list complete

        if (sp.probe & 1) continue;  // profiling: sort phases only
        // ---- walk the sorted list: every 8-lane group owns a contiguous range of entries ------------------------------------------
        const int E = cell_off[kMaxCells];
        const int e_lo = E * grp / kGroups, e_hi = E * (grp + 1) / kGroups;  // E <= 4608, grp < 48: fits 32 bits
        // grad_out row of query q_lo + q: gobase + q*pix floats; streamed from L2, prefetched two entries ahead
        const float *gobase = grad_out + ((n * Lq + q_lo) * sp.M + m) * (long long)kD + chf;
        unsigned grow_b = (unsigned)pix * 4u;  // bytes between the grad_out rows of consecutive queries
        pin(grow_b);
        // row (W+1) rows BEFORE pixel (0,0) of the level (never dereferenced there: see cinfo)
        float *gbase = grad_value + ((n * sp.S + sp.start[l]) - (W + 1)) * (long long)pix + (long long)m * kD + chf;
        const unsigned o10 = (unsigned)(W * pix);
        pin(gobase); pin(gbase);
        int cur = -1;
        const float2 z2 = make_float2(0.f, 0.f);
        f4p a00 = {z2, z2}, a01 = a00, a10 = a00, a11 = a00;  // sums of att * (1, lw, lh, lh*lw) * g over the current cell's samples
        auto flush = [&]() {
            if (sp.probe & 2) return;
            // (S0, S1, S2, S3) -> the four bilinear weights: (1-lh)(1-lw) = S0 - S1 - S2 + S3, (1-lh) lw = S1 - S3, lh (1-lw) = S2 - S3, lh lw = S3
            unsigned ci;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ci) : "r"(cinfo_s + cur * 4));
            float *gk = gbase + (ci & 0xfffff) * (unsigned)pix;  // the cell's top-left neighbour, pixel (cy-1, cx-1)
            const f4p r10 = psub(a10, a11), r01 = psub(a01, a11), r00 = psub(psub(a00, a01), r10);
            red_add_v4_if(ci & (1 << 20), gk, r00.lo.x, r00.lo.y, r00.hi.x, r00.hi.y);
            red_add_v4_if(ci & (2 << 20), gk + pix, r01.lo.x, r01.lo.y, r01.hi.x, r01.hi.y);
            gk += o10;
            red_add_v4_if(ci & (4 << 20), gk, r10.lo.x, r10.lo.y, r10.hi.x, r10.hi.y);
            red_add_v4_if(ci & (8 << 20), gk + pix, a11.lo.x, a11.lo.y, a11.hi.x, a11.hi.y);
        };
        auto issue_g = [&](int e2) -> float4 {
            unsigned packed;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(packed) : "r"(ent_s + e2 * 16 + 12));
            return ldg4(reinterpret_cast<const float *>(reinterpret_cast<const char *>(gobase) + (packed & 0xffff) * grow_b));
        };
        auto sample = [&](const float4 &en, const float4 &gr) {
            const int cell = __float_as_int(en.w) >> 16;
            if (cell != cur) {
                if (cur >= 0) flush();
                cur = cell;
                a00 = a01 = a10 = a11 = {z2, z2};
            }
            const float lh = en.x, lw = en.y, av = en.z;
            const float a1s = av * lw, a2s = av * lh, a3s = a2s * lw;
            const float2 b0 = make_float2(av, av), b1 = make_float2(a1s, a1s), b2 = make_float2(a2s, a2s), b3 = make_float2(a3s, a3s);
            const f4p g = pk(gr);
            a00.lo = __ffma2_rn(b0, g.lo, a00.lo); a00.hi = __ffma2_rn(b0, g.hi, a00.hi);
            a01.lo = __ffma2_rn(b1, g.lo, a01.lo); a01.hi = __ffma2_rn(b1, g.hi, a01.hi);
            a10.lo = __ffma2_rn(b2, g.lo, a10.lo); a10.hi = __ffma2_rn(b2, g.hi, a10.hi);
            a11.lo = __ffma2_rn(b3, g.lo, a11.lo); a11.hi = __ffma2_rn(b3, g.hi, a11.hi);
        };
        // software pipeline: grad_out rows are loaded TWO entries ahead into three rotating registers (the loop is unrolled by
        // three so the rotation costs no moves).  Reads past e_hi stay inside the list or hit the four zero entries behind it.
        if (e_hi > e_lo) {
            float4 g0 = issue_g(e_lo), g1 = issue_g(e_lo + 1), g2;
            int e = e_lo;
#pragma unroll 1
            for (; e + 3 <= e_hi; e += 3) {
                g2 = issue_g(e + 2);
                sample(lds4(ent_s + e * 16), g0);
                g0 = issue_g(e + 3);
                sample(lds4(ent_s + (e + 1) * 16), g1);
                g1 = issue_g(e + 4);
                sample(lds4(ent_s + (e + 2) * 16), g2);
            }
            if (e < e_hi) sample(lds4(ent_s + e * 16), g0);
            if (e + 1 < e_hi) sample(lds4(ent_s + (e + 1) * 16), g1);
            flush();
        }
    }
}

// (threads per CTA, CTAs per SM): the sorted list of a piece is the only shared memory, so several CTAs share an SM and the sort
// phases of one run under the walks of the others
struct Config { int threads, ctas; };
constexpr Config kConfigs[] = {{384, 3}, {512, 2}, {256, 3}, {288, 3}, {448, 2}, {384, 2}};
int scatter_config()
{
    static const int v = [] { const char *e = getenv("MSDA_SCATTER_CFG"); const int x = e ? atoi(e) : 0; return x >= 0 && x < 6 ? x : 0; }();  // A/B switch
    return v;
}

bool make_params(const Problem &p, ScatterParams &sp, size_t &smem)
{
    const int ctas = kConfigs[scatter_config()].ctas;
    if (!(p.dtype == 0 && p.loc_dtype == 0 && p.D == kD && p.have_host && p.L <= kMaxLevels)) return false;
    if (p.Lq < 1 || p.Lq > 65535 || p.P < 1 || p.P > 64) return false;
    sp.L = p.L; sp.P = p.P; sp.S = p.S; sp.M = p.M; sp.Lq = p.Lq;
    for (int l = 0; l < p.L; ++l) {
        sp.H[l] = p.table.H[l];
        sp.W[l] = p.table.W[l];
        sp.start[l] = p.table.start[l];
        sp.rw[l] = 1.0f / (float)sp.W[l];
        sp.rh[l] = 1.0f / (float)sp.H[l];
        int cs = 0;
        while ((1 << cs) < sp.W[l] + 1) ++cs;
        sp.cshift[l] = cs;
        if (((long long)(sp.H[l] + 1) << cs) > kMaxCells) return false;
        if ((long long)(sp.H[l] + 2) * (sp.W[l] + 2) * p.M * kD >= (1LL << 31)) return false;  // 32-bit row offsets
        if ((long long)(sp.H[l] + 2) * (sp.W[l] + 2) >= (1 << 20)) return false;                // 20-bit row offsets in cinfo
    }
    sp.total_q = (long long)p.N * p.M * p.L * p.Lq;
    if ((long long)p.N * p.M * p.L >= (1LL << 30)) return false;
    // a piece holds at most qmax*P sorted entries; `ctas` CTAs share an SM (227 KB + 1 KB reserved each)
    const size_t fixed = 4 * 16 + (kMaxCells + 8 + 32 + kMaxCells) * sizeof(int);
    const size_t budget = ((size_t)p.smem_optin + 1024) / ctas - 1024 - 512;
    if (budget < fixed + 64 * 16) return false;
    long long qmax = (long long)(budget - fixed) / 16 / p.P;
    if (qmax > p.Lq) qmax = p.Lq;
    if (qmax * p.P > 65535 - 4) qmax = (65535 - 4) / p.P;
    if (qmax < 1) return false;
    sp.qmax = (int)qmax;
    smem = (size_t)sp.qmax * p.P * 16 + fixed;
    sp.fused = p.ref != nullptr ? 1 : 0;
    static const int probe = [] { const char *e = getenv("MSDA_SCATTER_PROBE"); return e ? atoi(e) : 0; }();
    sp.probe = probe;
    return true;
}

template <auto Kern> cudaError_t ensure_smem_optin(const Problem &p)
{
    static std::atomic<unsigned long long> done{0};
    const unsigned long long bit = 1ULL << (p.device & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_optin);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

template <int CFG>
cudaError_t launch_cfg(const Problem &p, const BackwardIO &io, const ScatterParams &sp, size_t smem, cudaStream_t st)
{
    constexpr int T = kConfigs[CFG].threads, C = kConfigs[CFG].ctas;
    long long grid = (long long)p.sm_count * C;
    const long long pieces = (sp.total_q + sp.qmax - 1) / sp.qmax;
    if (grid > pieces) grid = pieces;
    auto go = [&](auto kern_c) {
        constexpr auto kern = decltype(kern_c)::value;
        cudaError_t e = ensure_smem_optin<kern>(p);
        if (e != cudaSuccess) return e;
        kern<<<(int)grid, T, smem, st>>>((const float *)p.loc, (const float *)p.attn, (const float *)io.grad_out, (float *)io.grad_value,
                                         (const float *)p.ref, sp);
        count_launch();
        return cudaGetLastError();
    };
    using std::integral_constant;
#define MSDA_SCATTER_KERN(PT, FU) integral_constant<decltype(&bwd_scatter32_kernel<PT, FU, T, C>), &bwd_scatter32_kernel<PT, FU, T, C>>()
    if (sp.fused) return p.P == 4 ? go(MSDA_SCATTER_KERN(4, true)) : go(MSDA_SCATTER_KERN(0, true));
    return p.P == 4 ? go(MSDA_SCATTER_KERN(4, false)) : go(MSDA_SCATTER_KERN(0, false));
#undef MSDA_SCATTER_KERN
}

}  // namespace

bool scatter32_supported(const Problem &p)
{
    ScatterParams sp;
    size_t smem;
    return make_params(p, sp, smem);
}

// grad_value += ... for every sample; the caller has zeroed grad_value
cudaError_t bwd_scatter32(const Problem &p, const BackwardIO &io, cudaStream_t st)
{
    ScatterParams sp;
    size_t smem;
    if (!make_params(p, sp, smem)) return cudaErrorInvalidValue;
    if (sp.total_q == 0) return cudaSuccess;
    switch (scatter_config()) {
    case 1: return launch_cfg<1>(p, io, sp, smem, st);
    case 2: return launch_cfg<2>(p, io, sp, smem, st);
    case 3: return launch_cfg<3>(p, io, sp, smem, st);
    case 4: return launch_cfg<4>(p, io, sp, smem, st);
    case 5: return launch_cfg<5>(p, io, sp, smem, st);
    default: return launch_cfg<0>(p, io, sp, smem, st);
    }
}

}  // namespace msda
