// msda_sorted32.cu -- D == 32, float32, geometry known on the host: SORT-AND-SEGMENT backward.
//
// Why (measured on B200, tools/microbench.cu -> profiles/): scattering grad_value with atomics costs one 128-byte
// row reduction per (sample, neighbour) = 15.8 M rows for the S4 shape; red.global.add.v4.f32 sustains 47 G rows/s
// (337 us), shared-memory float atomics are CAS loops (45 G rows/s, 351 us).  The reference pays 1256 us for it
// (ms_deform_im2col_cuda.cuh:306-408, scalar atomicAdd at :130-157).  Everything a sample contributes, however, is
// decided by the bilinear CELL it falls into (its top-left neighbour): samples of one cell share their four
// neighbour rows.  So each CTA takes one (frame n, head m, level l), counting-sorts that level's Lq*P samples by
// cell in shared memory, and then walks the sorted list with "owner" lane groups:
//   - the four value rows of the current cell live in registers (loaded once per cell, not once per sample),
//   - grad_value for the four rows accumulates in registers and leaves with ONE vector reduction per row per cell
//     visit (about 1.5 M row reductions instead of 15.8 M),
//   - per sample only ONE 128-byte row is read: its grad_output row, from a shared-memory slab of grad_out[n,:,m,:],
//   - grad_sampling_loc / grad_attn_weight are 4-lane shuffle reductions and plain stores.
// A lane group is 4 lanes x 8 channels; group parity picks which of its two 16-byte chunks a lane loads first, so
// the 8 lanes of every LDS.128 phase cover all 32 banks: conflict-free random row reads.
#include <cmath>

#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kD = 32;
constexpr int kThreads = 512;
constexpr int kGroups = kThreads / 4;
constexpr int kRounds = 9;          // samples per thread held in registers between the count and the fill pass
constexpr int kMaxCells = 1024;     // (H+1)*(W+1) per level, scanned by 512 threads x 2
constexpr int kInvalid = 0xfff;

struct SortParams {
    int L, P, S, M, Lq;
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    long long units;  // N*M*L
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}

__device__ __forceinline__ float4 lds4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

__global__ void __launch_bounds__(kThreads, 1)
bwd_sorted32_kernel(const float *__restrict__ value, const float *__restrict__ loc, const float *__restrict__ attn,
                    const float *__restrict__ grad_out, float *__restrict__ grad_value, float *__restrict__ grad_loc,
                    float *__restrict__ grad_attn, const __grid_constant__ SortParams sp)
{
    extern __shared__ float4 smem4[];
    const int Lq = sp.Lq, P = sp.P, EP = Lq * P, LP = sp.L * P;
    float4 *G4 = smem4;                       // grad_out[n, :, m, :]  -> Lq rows of 8 float4
    float4 *ent = smem4 + (size_t)Lq * 8;     // sorted samples: {lh, lw, attn, cell<<20 | p<<16 | q}; later their results
    int *cell_off = reinterpret_cast<int *>(ent + EP);   // kMaxCells + 1 counters -> exclusive offsets
    int *warp_tot = cell_off + kMaxCells + 8;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int grp = tid >> 2, sub = tid & 3, par = grp & 1;
    const int c0 = 2 * sub + par, c1 = 2 * sub + 1 - par;  // the two 16-byte chunks of a row this lane owns (load order)
    const unsigned G_s = (unsigned)__cvta_generic_to_shared(G4);
    const int pix = sp.M * kD;

    const long long u_lo = sp.units * blockIdx.x / gridDim.x, u_hi = sp.units * (blockIdx.x + 1) / gridDim.x;
    long long resident = -1;
    for (long long u = u_lo; u < u_hi; ++u) {
        const long long nm = u / sp.L;
        const int l = (int)(u - nm * sp.L);
        const long long n = nm / sp.M;
        const int m = (int)(nm - n * sp.M);
        const int H = sp.H[l], W = sp.W[l], CW = W + 1, cells = (H + 1) * CW;
        const float Hf = (float)H, Wf = (float)W;
        __syncthreads();  // previous unit is completely finished with ent / cell_off (and G if the slab changes)
        if (nm != resident) {
            const float *src = grad_out + (n * Lq * sp.M + m) * kD;
            for (int i = tid; i < Lq * 8; i += kThreads) cp_async16(G4 + i, src + (long long)(i >> 3) * pix + (i & 7) * 4);
            resident = nm;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int i = tid; i <= cells; i += kThreads) cell_off[i] = 0;
        __syncthreads();

        // ---- pass 1: cell of every sample of this level, rank inside the cell (shared-memory integer atomics) ----------
        float sx[kRounds], sy[kRounds], sa[kRounds];
        int cr[kRounds];  // cell << 16 | rank, or -1
        const long long meta0 = ((n * Lq) * sp.M + m) * (long long)LP + l * P;  // sample (q=0, p=0) of this (n, m, l)
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int i = tid + r * kThreads;
            cr[r] = -1;
            sx[r] = sy[r] = sa[r] = 0.f;
            if (i < EP) {
                const int q = i / P, p = i - q * P;
                const long long s = meta0 + (long long)q * sp.M * LP + p;
                const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + s);
                sx[r] = xy.x;
                sy[r] = xy.y;
                sa[r] = __ldg(attn + s);
                const float h = __fmaf_rn(xy.y, Hf, -0.5f), w = __fmaf_rn(xy.x, Wf, -0.5f);
                if (h > -1.f && w > -1.f && h < Hf && w < Wf) {
                    const int cell = ((int)floorf(h) + 1) * CW + (int)floorf(w) + 1;
                    cr[r] = (cell << 16) | atomicAdd(&cell_off[cell], 1);
                } else {  // sample outside the map: all three gradients are exactly zero (cuh:352-356)
                    *reinterpret_cast<float2 *>(grad_loc + 2 * s) = make_float2(0.f, 0.f);
                    grad_attn[s] = 0.f;
                }
            }
        }
        __syncthreads();

        // ---- exclusive scan of the cell counters (<= 1024 cells, two per thread) -----------------------------------------
        {
            const int i0 = 2 * tid, i1 = 2 * tid + 1;
            const int v0 = i0 < cells ? cell_off[i0] : 0, v1 = i1 < cells ? cell_off[i1] : 0;
            int incl = v0 + v1;
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
            const int base = (wid > 0 ? warp_tot[wid - 1] : 0) + incl - (v0 + v1);
            if (i0 < cells) cell_off[i0] = base;
            if (i1 < cells) cell_off[i1] = base + v0;
            if (tid == kThreads - 1) cell_off[kMaxCells] = warp_tot[kThreads / 32 - 1];  // number of in-range samples
        }
        __syncthreads();

        // ---- pass 2: write every in-range sample to its sorted slot ------------------------------------------------------
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            if (cr[r] >= 0) {
                const int i = tid + r * kThreads;
                const int q = i / P, p = i - q * P;
                const int cell = cr[r] >> 16, rank = cr[r] & 0xffff;
                const float h = __fmaf_rn(sy[r], Hf, -0.5f), w = __fmaf_rn(sx[r], Wf, -0.5f);
                const float lh = h - floorf(h), lw = w - floorf(w);
                ent[cell_off[cell] + rank] = make_float4(lh, lw, sa[r], __int_as_float((cell << 20) | (p << 16) | q));
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // sorted list complete, grad_out slab landed

        // ---- walk the sorted list: every 4-lane group owns a contiguous range of entries -----------------------------------
        const int E = cell_off[kMaxCells];
        const int e_lo = (int)((long long)E * grp / kGroups), e_hi = (int)((long long)E * (grp + 1) / kGroups);
        const int iters = (E + kGroups - 1) / kGroups;
        const long long lvl_row = (n * sp.S + sp.start[l]) * (long long)pix + (long long)m * kD;  // element offset of pixel (0,0), channel 0
        const long long out0 = ((n * Lq) * sp.M + m) * (long long)LP + l * P;
        int cur = -1;
        long long rowoff = 0;                 // element offset of the current cell's top-left neighbour (may be outside the map)
        bool in00 = false, in01 = false, in10 = false, in11 = false;
        float4 v00[2], dW[2], dH[2], dd[2];   // per owned chunk: V00, V01-V00, V10-V00, V11-V10-V01+V00
        float4 a00[2], a01[2], a10[2], a11[2];  // grad_value accumulators of the four neighbour rows
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2; ++k) a00[k] = a01[k] = a10[k] = a11[k] = z4;

        auto flush = [&]() {
            if (cur < 0) return;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int ch = (k == 0 ? c0 : c1) * 4;
                if (in00) red_add_v4(grad_value + rowoff + ch, a00[k].x, a00[k].y, a00[k].z, a00[k].w);
                if (in01) red_add_v4(grad_value + rowoff + pix + ch, a01[k].x, a01[k].y, a01[k].z, a01[k].w);
                if (in10) red_add_v4(grad_value + rowoff + (long long)W * pix + ch, a10[k].x, a10[k].y, a10[k].z, a10[k].w);
                if (in11) red_add_v4(grad_value + rowoff + (long long)(W + 1) * pix + ch, a11[k].x, a11[k].y, a11[k].z, a11[k].w);
                a00[k] = a01[k] = a10[k] = a11[k] = z4;
            }
        };

        for (int it = 0; it < iters; ++it) {
            const int e = e_lo + it;
            const bool active = e < e_hi;
            float s_a = 0.f, s_h = 0.f, s_w = 0.f;
            int q = 0, p = 0;
            float att = 0.f;
            if (active) {
                const float4 en = ent[e];
                const int packed = __float_as_int(en.w);
                const int cell = packed >> 20;
                q = packed & 0xffff;
                p = (packed >> 16) & 0xf;
                att = en.z;
                if (cell != cur) {
                    flush();
                    cur = cell;
                    const int cy = cell / CW, y0 = cy - 1, x0 = cell - cy * CW - 1;
                    const bool top = y0 >= 0, bot = y0 + 1 <= H - 1, lef = x0 >= 0, rig = x0 + 1 <= W - 1;
                    in00 = top && lef; in01 = top && rig; in10 = bot && lef; in11 = bot && rig;
                    rowoff = lvl_row + (long long)(y0 * W + x0) * pix;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int ch = (k == 0 ? c0 : c1) * 4;
                        const float4 t00 = in00 ? ldg4(value + rowoff + ch) : z4;
                        const float4 t01 = in01 ? ldg4(value + rowoff + pix + ch) : z4;
                        const float4 t10 = in10 ? ldg4(value + rowoff + (long long)W * pix + ch) : z4;
                        const float4 t11 = in11 ? ldg4(value + rowoff + (long long)(W + 1) * pix + ch) : z4;
                        v00[k] = t00;
                        dW[k] = f4sub(t01, t00);
                        dH[k] = f4sub(t10, t00);
                        dd[k] = f4sub(f4sub(t11, t10), dW[k]);
                    }
                }
                const float lh = en.x, lw = en.y;
                const float hh = 1.f - lh, hw = 1.f - lw;
                const float w1 = hh * hw * att, w2 = hh * lw * att, w3 = lh * hw * att, w4 = lh * lw * att;
                const unsigned grow = G_s + (unsigned)q * (kD * 4);
                const float4 g0 = lds4(grow + c0 * 16), g1 = lds4(grow + c1 * 16);
#define MSDA_CH(G, K, F)                                              \
    {                                                                 \
        const float gc = G.F;                                         \
        const float dh = fmaf(lw, dd[K].F, dH[K].F);                  \
        const float tp = fmaf(lw, dW[K].F, v00[K].F);                 \
        const float dw = fmaf(lh, dd[K].F, dW[K].F);                  \
        s_a = fmaf(gc, fmaf(lh, dh, tp), s_a);                        \
        s_h = fmaf(gc, dh, s_h);                                      \
        s_w = fmaf(gc, dw, s_w);                                      \
        a00[K].F = fmaf(w1, gc, a00[K].F);                            \
        a01[K].F = fmaf(w2, gc, a01[K].F);                            \
        a10[K].F = fmaf(w3, gc, a10[K].F);                            \
        a11[K].F = fmaf(w4, gc, a11[K].F);                            \
    }
                MSDA_CH(g0, 0, x) MSDA_CH(g0, 0, y) MSDA_CH(g0, 0, z) MSDA_CH(g0, 0, w)
                MSDA_CH(g1, 1, x) MSDA_CH(g1, 1, y) MSDA_CH(g1, 1, z) MSDA_CH(g1, 1, w)
#undef MSDA_CH
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                s_a += __shfl_xor_sync(0xffffffffu, s_a, o);
                s_h += __shfl_xor_sync(0xffffffffu, s_h, o);
                s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
            }
            if (active) {
                const long long s = out0 + (long long)q * sp.M * LP + p;
                if (sub == 0) *reinterpret_cast<float2 *>(grad_loc + 2 * s) = make_float2(Wf * att * s_w, Hf * att * s_h);
                if (sub == 1) grad_attn[s] = s_a;
            }
        }
        flush();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

bool make_params(const Problem &p, SortParams &sp, size_t &smem)
{
    if (!(p.dtype == 0 && p.loc_dtype == 0 && p.D == kD && p.have_host && p.L <= kMaxLevels)) return false;
    if (p.Lq < 1 || p.Lq > 65535 || p.P > 16 || (long long)p.Lq * p.P > (long long)kRounds * kThreads) return false;
    sp.L = p.L; sp.P = p.P; sp.S = p.S; sp.M = p.M; sp.Lq = p.Lq;
    for (int l = 0; l < p.L; ++l) {
        sp.H[l] = p.table.H[l];
        sp.W[l] = p.table.W[l];
        sp.start[l] = p.table.start[l];
        if ((long long)(sp.H[l] + 1) * (sp.W[l] + 1) > kMaxCells) return false;
    }
    sp.units = (long long)p.N * p.M * p.L;
    smem = (size_t)p.Lq * 8 * 16 + (size_t)p.Lq * p.P * 16 + (kMaxCells + 8 + 32) * sizeof(int);
    return smem <= (size_t)p.smem_optin;
}

}  // namespace

bool sorted32_supported(const Problem &p)
{
    SortParams sp;
    size_t smem;
    return make_params(p, sp, smem);
}

cudaError_t bwd_sorted32(const Problem &p, const BackwardIO &io, cudaStream_t st)
{
    SortParams sp;
    size_t smem;
    if (!make_params(p, sp, smem)) return cudaErrorInvalidValue;
    const long long n_val = (long long)p.N * p.S * p.M * p.D;
    if (n_val > 0) {
        cudaError_t e = cudaMemsetAsync(io.grad_value, 0, sizeof(float) * (size_t)n_val, st);
        if (e != cudaSuccess) return e;
    }
    if (sp.units == 0 || p.Lq == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(bwd_sorted32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (int)(sp.units < p.sm_count ? sp.units : p.sm_count);
    bwd_sorted32_kernel<<<grid, kThreads, smem, st>>>((const float *)p.value, (const float *)p.loc, (const float *)p.attn,
                                                      (const float *)io.grad_out, (float *)io.grad_value, (float *)io.grad_loc,
                                                      (float *)io.grad_attn, sp);
    count_launch();
    return cudaGetLastError();
}

}  // namespace msda
