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
//   - per sample only ONE 128-byte row is read: its grad_output row, streamed from L2 and prefetched one sample ahead
//     (the value level slab, zero-padded, sits in shared memory so a cell switch costs 8 LDS, not 8 L2 round trips),
//   - grad_sampling_loc / grad_attn_weight are 4-lane shuffle reductions and plain stores.
// MODE bit 0 (default on): the grad_out chunks do not travel through prefetch registers -- ptxas puts all of those LDGs on
// one scoreboard, so the first use of any of them waits for ALL of them, including the one issued a few instructions
// earlier (20 % long-scoreboard stalls in profiles/r01p) -- but through a private shared-memory slot per lane, filled by
// cp.async two steps ahead and read back after cp.async.wait_group 2: a counted, in-order wait.
// MODE bit 1: the reduction + stores of sample i are deferred into the basic block of sample i+1.
// Every unit also prefetches the NEXT unit's locations, weights and value rows (and the next slab's grad_out rows) into L2.
// A lane group is 8 lanes x 4 channels (one 128-byte row per LDS.128 phase: conflict-free); the three per-sample sums
// are reduced over the 8 lanes with a transposed butterfly (4 shuffles instead of 9).
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kD = 32;
constexpr int kMaxSamples = 4608;    // Lq*P per level that the register-resident count/fill passes can hold
//         // samples per thread held in registers between the count and the fill pass
constexpr int kMaxCells = 1024;     // cell-grid size per level, scanned two per thread

struct SortParams {
    int L, P, S, M, Lq;
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    int cshift[kMaxLevels];  // log2 of the cell-grid pitch (power of two >= W+1): cell = (y0+1) << cshift | (x0+1)
    unsigned wrecip[kMaxLevels];  // ceil(2^32 / W): pixel / W == umulhi(pixel, wrecip) for pixel < 2^16
    int vrows_max;           // rows of the largest zero-padded level slab, (H+2)*(W+2)
    long long units;         // N*M*L
    int base;                // whole units per CTA
    int rem;                 // leftover units (units - base*grid), each cut into `parts` pieces: one piece per CTA
    int parts;
    int zero_frames;         // whole units of frames < zero_frames zero their own grad_value slab (the host zeroes the other frames)
    int prefetch;            // L2-prefetch the next unit's inputs during the walk
    int fused;               // `loc` holds raw offsets: location = ref + offset / (W, H); grad_loc is then d/d(offset)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}


// packed fp32x2 helpers (FFMA2 / FADD2 on sm_100a): halve the issue slots of the per-channel math
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

// Pin a value in a register: ptxas otherwise re-derives "cheap" values (special registers, parameter loads, the
// shared-window base) inside the hot loop, paying an issue slot each time.
__device__ __forceinline__ void pin(int &x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(unsigned &x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(float &x) { asm volatile("" : "+f"(x)); }
template <typename T> __device__ __forceinline__ void pin(T *&x) { asm volatile("" : "+l"(x)); }
__device__ __forceinline__ float4 lds4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__host__ __device__ constexpr int threads_for_split(int split) { return split == 8 ? 768 : split == 4 ? 512 : 384; }

using bf16 = __nv_bfloat16;
__device__ __forceinline__ float4 load4(const float *p) { return ldg4(p); }
__device__ __forceinline__ float4 load4(const bf16 *p)
{
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float2 load2(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ float2 load2(const bf16 *p) { return __bfloat1622float2(__ldg(reinterpret_cast<const __nv_bfloat162 *>(p))); }
__device__ __forceinline__ float load1(const float *p) { return __ldg(p); }
__device__ __forceinline__ float load1(const bf16 *p) { return __bfloat162float(__ldg(p)); }
__device__ __forceinline__ void store1(float *p, float v) { *p = v; }
__device__ __forceinline__ void store1(bf16 *p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store2(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(bf16 *p, float a, float b) { *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b); }

// SPLIT lanes share one sample; each lane owns CH = 8/SPLIT chunks of 4 channels.
// TV: value / grad_out element type (float or bf16; grad_value always accumulates in fp32: for bf16 into a workspace that a
// second kernel narrows).  TM: element type of sampling locations, attention weights and their gradients.
template <typename TV, typename TM, int P_T, int SPLIT, int MODE>
__global__ void __launch_bounds__(threads_for_split(SPLIT), 1)
bwd_sorted32_kernel(const TV *__restrict__ value, const TM *__restrict__ loc, const TM *__restrict__ attn,
                    const TV *__restrict__ grad_out, float *__restrict__ grad_value, TM *__restrict__ grad_loc,
                    TM *__restrict__ grad_attn, const TM *__restrict__ ref, const unsigned char *__restrict__ mask,
                    const __grid_constant__ SortParams sp)
{
    constexpr int kThreads = threads_for_split(SPLIT), CH = 8 / SPLIT;
    constexpr int kGroups = kThreads / SPLIT, kRounds = (kMaxSamples + kThreads - 1) / kThreads;
    constexpr int kCPT = (kMaxCells + kThreads - 1) / kThreads;  // cells scanned per thread
    extern __shared__ float4 smem4[];
    const int Lq = sp.Lq, P = P_T > 0 ? P_T : sp.P, EP = Lq * P, LP = sp.L * P;
    float4 *V4 = smem4;                                   // zero-padded value slab of the current level: (H+2) x (W+2) rows
    float4 *ent = smem4 + (size_t)sp.vrows_max * 8;       // sorted samples: {lh, lw, attn, cell<<20 | p<<16 | q}
    int *cell_off = reinterpret_cast<int *>(ent + EP + 4);  // (four spare entries, see the prefetch) kMaxCells + 1 counters -> exclusive offsets
    int *warp_tot = cell_off + kMaxCells + 8;
    constexpr bool GRING = (MODE & 1) != 0, DEFER = (MODE & 2) != 0, FUSED = (MODE & 4) != 0;
    // MODE bit 4 (16): fewer FMA-pipe instructions per sample -- four dot products of grad_out with (V00, dW, dH, dd) combined after
    // the horizontal add, and grad_value accumulated in the basis (1, lw, lh, lh*lw) and converted to the four neighbours at flush
    constexpr bool LEAN = (MODE & 16) != 0;
    // GRING: slot (stage, k) of thread t at gring + ((stage*CH + k) * kThreads + t) * 16 bytes, three stages
    // per-cell facts of the current level, filled while the unit is sorted: {value-slab byte offset of the cell's top-left
    // neighbour | 4-bit mask of the neighbours inside the map, grad_value element offset of that neighbour}
    int2 *cinfo = reinterpret_cast<int2 *>(warp_tot + 32);
    unsigned cinfo_s = (unsigned)__cvta_generic_to_shared(cinfo);
    pin(cinfo_s);
    unsigned gslot_s = (unsigned)__cvta_generic_to_shared(cinfo + kMaxCells) + threadIdx.x * 16;
    pin(gslot_s);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int grp = tid / SPLIT;
    int j = tid % SPLIT;                                  // lane j of its sample group
    pin(j);
    // slot k of this lane is chunk j*CH + ((k + grp) % CH): the 8 lanes of an LDS.128 phase then cover all 32 banks
    int chf[CH];                                          // float offset of slot k inside a 32-channel row
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        chf[k] = (j * CH + ((k + grp) % CH)) * 4;
        pin(chf[k]);
    }
    unsigned ent_s = (unsigned)__cvta_generic_to_shared(ent);
    pin(ent_s);
    const int pix = sp.M * kD;
    const int qstride = sp.M * LP;  // samples between consecutive queries of one head

    // schedule: CTA b runs whole units [b*base, (b+1)*base) and then, if it exists, piece b of the leftover units
    // (leftover unit grid*base + b/parts, entry range part b%parts of `parts`): without the split the CTAs that get
    // one unit more than the others set the kernel time (960 units over 148 CTAs: 7 vs 6.49).
    const int n_mine = sp.base + ((int)blockIdx.x < sp.rem * sp.parts ? 1 : 0);
    for (int ui = 0; ui < n_mine; ++ui) {
        const bool piece = ui == sp.base;
        // (32-bit unit arithmetic: every thread decodes its unit, and 64-bit divisions are ~100 instructions each)
        const int u = piece ? (int)gridDim.x * sp.base + (int)blockIdx.x / sp.parts : (int)blockIdx.x * sp.base + ui;
        const int part = piece ? (int)blockIdx.x % sp.parts : 0, parts = piece ? sp.parts : 1;
        const int nm = u / sp.L;
        const int l = u - nm * sp.L;
        const long long n = nm / sp.M;
        const int m = nm - (int)n * sp.M;
        const int H = sp.H[l], W = sp.W[l], cs = sp.cshift[l], cells = (H + 1) << cs, Wp = W + 2;
        const float Hf = (float)H, Wf = (float)W;
        const long long lvl_row = (n * sp.S + sp.start[l]) * (long long)pix + (long long)m * kD;  // pixel (0,0), channel 0
        const TV *vbase = value + lvl_row;
        __syncthreads();  // previous unit is completely finished with V / ent / cell_off
        // ---- pass 1: cell of every sample of this level, rank inside the cell (shared-memory integer atomics) ----------
        // sample (q, p) of this (n, m, l) sits at loc2[q*qstride + p]; per-unit base pointers keep the index math 32-bit
        const long long meta0 = ((n * Lq) * sp.M + m) * (long long)LP + l * P;
        const TM *locb = loc + 2 * meta0;
        const TM *att = attn + meta0;
        TM *gloc = grad_loc + 2 * meta0;
        TM *gatt = grad_attn + meta0;
        float sx[kRounds], sy[kRounds], sa[kRounds];
        int cr[kRounds];  // cell << 16 | rank, or -1
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {  // all loads first: one L2 round trip for the whole pass
            const int i = tid + r * kThreads;
            sx[r] = sy[r] = 2.f;  // out of range
            sa[r] = 0.f;
            if (i < EP) {
                const int q = P_T > 0 ? i / P_T : i / P, p = i - q * P;
                float2 xy = load2(locb + 2 * (q * qstride + p));
                if (FUSED) {  // module prologue (ms_deform_attn.py:110-112): reference point + offset / (W_l, H_l)
                    const float2 rf = load2(ref + ((n * Lq + q) * sp.L + l) * 2);
                    // (the same FMA with the host-rounded reciprocal as the forward: a sample lands in the same cell in both passes)
                    xy.x = __fmaf_rn(xy.x, 1.0f / Wf, rf.x);
                    xy.y = __fmaf_rn(xy.y, 1.0f / Hf, rf.y);
                }
                sx[r] = xy.x;
                sy[r] = xy.y;
                sa[r] = load1(att + q * qstride + p);
            }
        }
        // (the slab copy and the zero fills below overlap the latency of the loads just issued)
        // value level slab -> shared memory (interior by cp.async, the one-pixel border zeroed by hand)
        if (sizeof(TV) == 4) {
            for (int i = tid; i < H * W * 8; i += kThreads) {
                const int pxl = i >> 3, c = i & 7;
                const int y = (int)__umulhi((unsigned)pxl, sp.wrecip[l]), x = pxl - y * W;
                cp_async16(V4 + ((y + 1) * Wp + x + 1) * 8 + c, vbase + pxl * pix + c * 4);
            }
        } else {  // bf16 rows (64 bytes) are widened to fp32 on the way in
            for (int i = tid; i < H * W * 8; i += kThreads) {
                const int pxl = i >> 3, c = i & 7;
                const int y = (int)__umulhi((unsigned)pxl, sp.wrecip[l]), x = pxl - y * W;
                V4[((y + 1) * Wp + x + 1) * 8 + c] = load4(vbase + pxl * pix + c * 4);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // A unit this CTA owns alone zeroes its own slab of grad_value here, right before it accumulates into it: the lines are
        // still in L2 when the reductions arrive, and no 42 MB memset runs ahead of the kernel.  (The few leftover units that
        // several CTAs share lie in the frames >= zero_frames, which the host zeroes with a small memset.)
        if (!piece && n < sp.zero_frames) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            float *gvl = grad_value + lvl_row;
            for (int i = tid; i < H * W * 8; i += kThreads) *reinterpret_cast<float4 *>(gvl + (i >> 3) * pix + (i & 7) * 4) = z;
        }
        {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < Wp * 8; i += kThreads) {  // top and bottom padded rows
                V4[i] = z;
                V4[(H + 1) * Wp * 8 + i] = z;
            }
            for (int i = tid; i < H * 16; i += kThreads) {  // left and right padded columns
                const int y = (i >> 4) + 1, side = (i >> 3) & 1, c = i & 7;
                V4[(y * Wp + side * (W + 1)) * 8 + c] = z;
            }
        }
        for (int i = tid; i <= cells; i += kThreads) cell_off[i] = 0;
        for (int c = tid; c < cells; c += kThreads) {
            const int cy = c >> cs, cx = c & ((1 << cs) - 1);
            const int inm = ((cy >= 1 && cx >= 1) ? 1 : 0) | ((cy >= 1 && cx <= W - 1) ? 2 : 0) |
                            ((cy <= H - 1 && cx >= 1) ? 4 : 0) | ((cy <= H - 1 && cx <= W - 1) ? 8 : 0);
            cinfo[c] = make_int2((cy * Wp + cx) * (kD * 4) | inm, ((cy - 1) * W + cx - 1) * pix);
        }

        __syncthreads();  // counters are zero
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int i = tid + r * kThreads;
            cr[r] = -1;
            if (i < EP) {
                const float h = __fmaf_rn(sy[r], Hf, -0.5f), w = __fmaf_rn(sx[r], Wf, -0.5f);
                if (h > -1.f && w > -1.f && h < Hf && w < Wf) {
                    int y0, x0;
                    floor_magic(h, y0);
                    floor_magic(w, x0);
                    const int cell = ((y0 + 1) << cs) + x0 + 1;
                    cr[r] = (cell << 16) | atomicAdd(&cell_off[cell], 1);
                } else if (part == 0) {  // sample outside the map: all three gradients are exactly zero (cuh:352-356)
                    const int q = P_T > 0 ? i / P_T : i / P, p = i - q * P;
                    const int s = q * qstride + p;
                    store2(gloc + 2 * s, 0.f, 0.f);
                    store1(gatt + s, 0.f);
                }
            }
        }
        __syncthreads();

        // ---- exclusive scan of the cell counters (<= kMaxCells cells, kCPT consecutive cells per thread) -------------------
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
                // the grad_out prefetch runs up to four entries past a group's range without clamping: past the list's end it must
                // still find a valid query index (0)
                for (int k = 0; k < 4; ++k) ent[tot + k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();

        // ---- pass 2: write every in-range sample to its sorted slot ------------------------------------------------------
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            if (cr[r] >= 0) {
                const int i = tid + r * kThreads;
                const int q = P_T > 0 ? i / P_T : i / P, p = i - q * P;
                const int cell = cr[r] >> 16, rank = cr[r] & 0xffff;
                const float h = __fmaf_rn(sy[r], Hf, -0.5f), w = __fmaf_rn(sx[r], Wf, -0.5f);
                int dummy;
                const float lh = h - floor_magic(h, dummy), lw = w - floor_magic(w, dummy);
                ent[cell_off[cell] + rank] = make_float4(lh, lw, sa[r], __int_as_float((cell << 20) | (p << 16) | q));
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (FUSED && mask != nullptr) {  // padding mask folded in: masked pixels read as zero rows (their grad_value is zeroed afterwards)
            __syncthreads();
            const unsigned char *mk = mask + n * sp.S + sp.start[l];
            for (int pxl = tid; pxl < H * W; pxl += kThreads) {
                if (mk[pxl]) {
                    const int y = (int)__umulhi((unsigned)pxl, sp.wrecip[l]), x = pxl - y * W;
#pragma unroll
                    for (int c = 0; c < 8; ++c) V4[((y + 1) * Wp + x + 1) * 8 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        __syncthreads();  // sorted list complete, value slab landed

        // ---- L2 prefetch of the next whole unit of this CTA: its locations / weights, its value rows and, when it starts a new
        // (frame, head) slab, its grad_out rows -- so that the next sort and walk find them in L2 instead of paying HBM latency
        if (sp.prefetch && ui + 1 < sp.base) {
            const int u2 = u + 1, nm2 = u2 / sp.L;
            const int l2 = u2 - nm2 * sp.L;
            const long long n2 = nm2 / sp.M;
            const int m2 = nm2 - (int)n2 * sp.M;
            const long long meta2 = ((n2 * Lq) * sp.M + m2) * (long long)LP + l2 * P;
            const TM *loc2 = loc + 2 * meta2, *att2 = attn + meta2;
            for (int q = tid; q < Lq; q += kThreads) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(loc2 + 2 * (q * qstride)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(att2 + q * qstride));
            }
            const TV *v2 = value + (n2 * sp.S + sp.start[l2]) * (long long)pix + (long long)m2 * kD;
            if (sp.prefetch == 1)
                for (int i = tid; i < sp.H[l2] * sp.W[l2]; i += kThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(v2 + i * pix));
            if (sp.prefetch == 1 && l2 == 0) {
                const TV *g2p = grad_out + (n2 * Lq * sp.M + m2) * (long long)kD;
                for (int q = tid; q < Lq; q += kThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(g2p + q * pix));
            }
        }
        // ---- walk the sorted list: every SPLIT-lane group owns a contiguous range of entries -------------------------------
        const int Etot = cell_off[kMaxCells];
        // a piece takes whole cells (the order INSIDE a cell differs between the CTAs that sort the same unit, the cell offsets do
        // not): its range starts at the first cell whose offset reaches Etot*part/parts
        auto cell_boundary = [&](int target) {
            int lo = 0, hi = cells;  // first c in [0, cells] with cell_off[c] >= target (cell_off[cells] stands for Etot)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cell_off[mid] >= target) hi = mid; else lo = mid + 1;
            }
            return lo < cells ? cell_off[lo] : Etot;
        };
        int E0 = 0, E = Etot;
        if (parts > 1) {
            E0 = part == 0 ? 0 : cell_boundary(Etot * part / parts);
            E = (part + 1 == parts ? Etot : cell_boundary(Etot * (part + 1) / parts)) - E0;
        }
        const int e_lo = E0 + E * grp / kGroups, e_hi = E0 + E * (grp + 1) / kGroups;  // E <= 4608, grp < 384: fits 32 bits
        // grad_out row of query q: gobase + q*pix (+ chf[k]); streamed from L2, prefetched one entry ahead
        const TV *gobase = grad_out + (n * Lq * sp.M + m) * (long long)kD + chf[0];  // (this lane's first chunk folded in)
        unsigned grow_b = (unsigned)pix * (unsigned)sizeof(TV);  // bytes between the grad_out rows of consecutive queries
        pin(grow_b);
        float *gbase = grad_value + lvl_row;
        const int o01 = pix, o10 = W * pix, o11 = (W + 1) * pix;
        unsigned V_s = (unsigned)__cvta_generic_to_shared(V4);
        pin(V_s);
        int qs = qstride, Wpb = Wp * (kD * 4);
        pin(qs); pin(Wpb);
        // result stores: SPLIT >= 4: lane 0 grad_attn, lane 1 grad_loc.y, lane 2 grad_loc.x; SPLIT == 2: lane 0 grad_loc (float2), lane 1 grad_attn
        TM *res_base = SPLIT >= 4 ? (j == 0 ? gatt : gloc + (j == 1 ? 1 : 0)) : (j == 0 ? gloc : gatt);
        int res_stride = SPLIT >= 4 ? (j == 0 ? 1 : 2) : (j == 0 ? 2 : 1);
        // d(loc)/d(offset) = 1 / (W, H) cancels the (W, H) factor of grad_sampling_loc (cuh:126-127) in the fused form
        float res_scale = (j == 0 || FUSED) ? 1.f : (j == 1 ? Hf : Wf);
        pin(res_base); pin(res_stride); pin(res_scale); pin(gobase); pin(gbase);
        int cur = -1, inmask = 0;
        float *gp = gbase;             // grad_value row of the current cell's top-left neighbour
        f4p v00[CH], dW[CH], dH[CH], dd[CH];       // V00, V01-V00, V10-V00, V11-V10-V01+V00 per owned chunk
        f4p a00[CH], a01[CH], a10[CH], a11[CH];    // grad_value accumulators of the four neighbour rows
        const float2 z2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CH; ++k) a00[k] = a01[k] = a10[k] = a11[k] = {z2, z2};

        auto flush = [&]() {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                float *gk = gp + chf[k];
                f4p r00 = a00[k], r01 = a01[k], r10 = a10[k];
                if (LEAN) {  // (S0, S1, S2, S3) = sums of att * (1, lw, lh, lh*lw) * g  ->  the four bilinear weights (into temporaries:
                             // the accumulators themselves are only ever zeroed, so no register moves at the branch merge)
                    r10 = psub(a10[k], a11[k]);                    // lh*(1-lw)
                    r00 = psub(psub(a00[k], a01[k]), r10);         // (1-lh)*(1-lw) = S0 - S1 - (S2 - S3)
                    r01 = psub(a01[k], a11[k]);                    // (1-lh)*lw
                }
                red_add_v4_if(inmask & 1, gk, r00.lo.x, r00.lo.y, r00.hi.x, r00.hi.y);
                red_add_v4_if(inmask & 2, gk + o01, r01.lo.x, r01.lo.y, r01.hi.x, r01.hi.y);
                red_add_v4_if(inmask & 4, gk + o10, r10.lo.x, r10.lo.y, r10.hi.x, r10.hi.y);
                red_add_v4_if(inmask & 8, gk + o11, a11[k].lo.x, a11[k].lo.y, a11[k].hi.x, a11[k].hi.y);
            }
        };
        // start the loads of the grad_out chunks of sorted entry e2 (clamped to this group's range): into the lane's shared-memory
        // slot `si` (GRING; one cp.async group per call) or into the register set g
        auto issue_g = [&](int e2, int si, float4 (&g)[CH]) {
            unsigned packed;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(packed) : "r"(ent_s + e2 * 16 + 12));
            // 64-bit base + unsigned 32-bit byte offset: two adds instead of a sign-extended 64-bit multiply-shift chain
            const TV *gq = reinterpret_cast<const TV *>(reinterpret_cast<const char *>(gobase) + (packed & 0xffff) * grow_b);
            if (GRING) {
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const unsigned slot = gslot_s + (unsigned)((si * CH + k) * kThreads) * 16;
                    if (sizeof(TV) == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(gq + (chf[k] - chf[0])) : "memory");
                    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(slot), "l"(gq + (chf[k] - chf[0])) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            } else {
#pragma unroll
                for (int k = 0; k < CH; ++k) g[k] = load4(gq + (chf[k] - chf[0]));
            }
        };
        // GRING: wait for the copy into slot `si` (all but the two newest groups, or everything when `drain`) and read it back
        auto fetch_g = [&](int si, bool drain, float4 (&g)[CH]) {
            if (!GRING) return;
            if (drain) asm volatile("cp.async.wait_group 0;" ::: "memory");
            else asm volatile("cp.async.wait_group 2;" ::: "memory");
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const unsigned slot = gslot_s + (unsigned)((si * CH + k) * kThreads) * 16;
                if (sizeof(TV) == 4) {
                    g[k] = lds4(slot);
                } else {
                    uint2 r;
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(slot));
                    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
                    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
                    g[k] = make_float4(a.x, a.y, b.x, b.y);
                }
            }
        };
        // deferred tail of the previous sample (DEFER): its three lane-partial sums, weight and result index
        float d_a = 0.f, d_h = 0.f, d_w = 0.f;
        int d_idx = -1;
        // s_h / s_w arrive already multiplied by the sample's attention weight (cuh:126-127), so the tail is one multiply per lane
        auto finish = [&](float s_a, float s_h, float s_w, int sidx) {  // SPLIT >= 4: transposed butterfly + stores
            float x = (j & 1) ? s_h : s_a;
            x += __shfl_xor_sync(0xffffffffu, (j & 1) ? s_a : s_h, 1);
            s_w += __shfl_xor_sync(0xffffffffu, s_w, 1);
            float keep = (j & 2) ? s_w : x;
            keep += __shfl_xor_sync(0xffffffffu, (j & 2) ? x : s_w, 2);
            if (SPLIT == 8) keep += __shfl_xor_sync(0xffffffffu, keep, 4);
            if (sidx >= 0 && j < 3)
                store1(res_base + sidx * res_stride, keep * res_scale);
        };
        // one sample: en = its sorted entry, gr = its grad_out chunks (both already in registers)
#define MSDA_SAMPLE_BODY \
            float s_a = 0.f, s_h = 0.f, s_w = 0.f, att_v = 0.f; \
            int sidx = 0; \
            if (active) { \
                const int packed = __float_as_int(en.w); \
                const int cell = packed >> 20; \
                sidx = (packed & 0xffff) * qs + ((packed >> 16) & 0xf); \
                att_v = en.z; \
                if (cell != cur) { \
                    flush(); \
                    cur = cell; \
                    int2 ci; \
                    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(ci.x), "=r"(ci.y) : "r"(cinfo_s + cell * 8)); \
                    inmask = ci.x & 15; \
                    gp = gbase + ci.y; \
                    const unsigned vr = V_s + (unsigned)(ci.x & ~127); \
_Pragma("unroll") \
                    for (int k = 0; k < CH; ++k) { \
                        const unsigned a = vr + chf[k] * 4; \
                        const f4p t00 = pk(lds4(a)), t01 = pk(lds4(a + kD * 4)), t10 = pk(lds4(a + Wpb)), t11 = pk(lds4(a + Wpb + kD * 4)); \
                        v00[k] = t00; \
                        dW[k] = psub(t01, t00); \
                        dH[k] = psub(t10, t00); \
                        dd[k] = psub(psub(t11, t10), dW[k]); \
                        a00[k] = a01[k] = a10[k] = a11[k] = {z2, z2}; \
                    } \
                } \
                const float lh = en.x, lw = en.y; \
                if (LEAN) { \
                    const float a1s = att_v * lw, a2s = att_v * lh, a3s = a2s * lw; \
                    const float2 b0 = make_float2(att_v, att_v), b1 = make_float2(a1s, a1s), b2 = make_float2(a2s, a2s), b3 = make_float2(a3s, a3s); \
                    float2 qA = z2, qB = z2, qC = z2, qE = z2; \
_Pragma("unroll") \
                    for (int k = 0; k < CH; ++k) { \
                        const f4p g = pk(gr[k]); \
                        qA = __ffma2_rn(g.hi, v00[k].hi, __ffma2_rn(g.lo, v00[k].lo, qA)); \
                        qB = __ffma2_rn(g.hi, dW[k].hi, __ffma2_rn(g.lo, dW[k].lo, qB)); \
                        qC = __ffma2_rn(g.hi, dH[k].hi, __ffma2_rn(g.lo, dH[k].lo, qC)); \
                        qE = __ffma2_rn(g.hi, dd[k].hi, __ffma2_rn(g.lo, dd[k].lo, qE)); \
                        a00[k].lo = __ffma2_rn(b0, g.lo, a00[k].lo); a00[k].hi = __ffma2_rn(b0, g.hi, a00[k].hi); \
                        a01[k].lo = __ffma2_rn(b1, g.lo, a01[k].lo); a01[k].hi = __ffma2_rn(b1, g.hi, a01[k].hi); \
                        a10[k].lo = __ffma2_rn(b2, g.lo, a10[k].lo); a10[k].hi = __ffma2_rn(b2, g.hi, a10[k].hi); \
                        a11[k].lo = __ffma2_rn(b3, g.lo, a11[k].lo); a11[k].hi = __ffma2_rn(b3, g.hi, a11[k].hi); \
                    } \
                    const float A = qA.x + qA.y, B = qB.x + qB.y, C = qC.x + qC.y, E = qE.x + qE.y; \
                    s_w = __fmaf_rn(lh, E, B); \
                    s_h = __fmaf_rn(lw, E, C); \
                    s_a = __fmaf_rn(lh, s_h, __fmaf_rn(lw, B, A)); \
                } else { \
                const float ah = (1.f - lh) * att_v, al = lh * att_v, hw = 1.f - lw; \
                const float2 lh2 = make_float2(lh, lh), lw2 = make_float2(lw, lw); \
                const float2 w1 = make_float2(ah * hw, ah * hw), w2 = make_float2(ah * lw, ah * lw); \
                const float2 w3 = make_float2(al * hw, al * hw), w4 = make_float2(al * lw, al * lw); \
                float2 pa = z2, ph = z2, pw = z2; \
_Pragma("unroll") \
                for (int k = 0; k < CH; ++k) { \
                    const f4p g = pk(gr[k]); \
                    const float2 dh0 = __ffma2_rn(lw2, dd[k].lo, dH[k].lo), dh1 = __ffma2_rn(lw2, dd[k].hi, dH[k].hi); \
                    const float2 tp0 = __ffma2_rn(lw2, dW[k].lo, v00[k].lo), tp1 = __ffma2_rn(lw2, dW[k].hi, v00[k].hi); \
                    const float2 dw0 = __ffma2_rn(lh2, dd[k].lo, dW[k].lo), dw1 = __ffma2_rn(lh2, dd[k].hi, dW[k].hi); \
                    pa = __ffma2_rn(g.hi, __ffma2_rn(lh2, dh1, tp1), __ffma2_rn(g.lo, __ffma2_rn(lh2, dh0, tp0), pa)); \
                    ph = __ffma2_rn(g.hi, dh1, __ffma2_rn(g.lo, dh0, ph)); \
                    pw = __ffma2_rn(g.hi, dw1, __ffma2_rn(g.lo, dw0, pw)); \
                    a00[k].lo = __ffma2_rn(w1, g.lo, a00[k].lo); a00[k].hi = __ffma2_rn(w1, g.hi, a00[k].hi); \
                    a01[k].lo = __ffma2_rn(w2, g.lo, a01[k].lo); a01[k].hi = __ffma2_rn(w2, g.hi, a01[k].hi); \
                    a10[k].lo = __ffma2_rn(w3, g.lo, a10[k].lo); a10[k].hi = __ffma2_rn(w3, g.hi, a10[k].hi); \
                    a11[k].lo = __ffma2_rn(w4, g.lo, a11[k].lo); a11[k].hi = __ffma2_rn(w4, g.hi, a11[k].hi); \
                } \
                s_a = pa.x + pa.y; \
                s_h = ph.x + ph.y; \
                s_w = pw.x + pw.y; \
                } \
            } \
            if (SPLIT == 2) { \
                s_a += __shfl_xor_sync(0xffffffffu, s_a, 1); \
                s_h += __shfl_xor_sync(0xffffffffu, s_h, 1); \
                s_w += __shfl_xor_sync(0xffffffffu, s_w, 1); \
                if (active) { \
                    if (j == 0) store2(res_base + sidx * 2, (FUSED ? 1.f : Wf) * att_v * s_w, (FUSED ? 1.f : Hf) * att_v * s_h); \
                    else store1(res_base + sidx, s_a); \
                } \
            } else if (DEFER) { \
                finish(d_a, d_h, d_w, d_idx); \
                d_a = s_a; d_h = s_h * att_v; d_w = s_w * att_v; d_idx = active ? sidx : -1; \
            } else { \
                finish(s_a, s_h * att_v, s_w * att_v, active ? sidx : -1); \
            }
        // nvcc 12.9 crashes on a generic lambda here, hence two plain lambdas sharing one body
        auto sample_steady = [&](const float4 &en, const float4 (&gr)[CH]) {
            const bool active = true;  // steady-state steps are active for every group
            MSDA_SAMPLE_BODY
        };
        auto sample_checked = [&](const float4 &en, const float4 (&gr)[CH], bool active) {
            MSDA_SAMPLE_BODY
        };
#undef MSDA_SAMPLE_BODY

        // software pipeline: grad_out rows are loaded TWO samples ahead into three rotating register sets (the loop is unrolled
        // by three so the rotation costs no moves); one sample ahead does not cover the L2 latency.  Every group owns `full`
        // or `full`+1 entries, so the first `full` steps need no activity checks; the remainder is peeled.
        const int full = E / kGroups, cnt = e_hi - e_lo;
        const bool has_tail = full * kGroups != E;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 g0[CH], g1[CH], g2[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) g0[k] = g1[k] = g2[k] = z4;
        if (cnt > 0) {
            issue_g(e_lo, 0, g0);
            issue_g(e_lo + 1, 1, g1);
        }
        int it = 0;
        for (; it + 3 <= full; it += 3) {
            const int e = e_lo + it;
            issue_g(e + 2, 2, g2);
            fetch_g(0, false, g0);
            sample_steady(lds4(ent_s + e * 16), g0);
            issue_g(e + 3, 0, g0);
            fetch_g(1, false, g1);
            sample_steady(lds4(ent_s + (e + 1) * 16), g1);
            issue_g(e + 4, 1, g1);
            fetch_g(2, false, g2);
            sample_steady(lds4(ent_s + (e + 2) * 16), g2);
        }
        const int rem = full - it + (has_tail ? 1 : 0);  // 0..3 remaining steps, the same for every group
        if (rem >= 1) {
            const int e = e_lo + it;
            if (cnt > 0 && rem >= 3) issue_g(e + 2, 2, g2);
            fetch_g(0, true, g0);  // (the remainder drains the copy queue: its group count depends on rem)
            sample_checked(e < e_hi ? lds4(ent_s + e * 16) : z4, g0, e < e_hi);
        }
        if (rem >= 2) {
            const int e = e_lo + it + 1;
            fetch_g(1, true, g1);
            sample_checked(e < e_hi ? lds4(ent_s + e * 16) : z4, g1, e < e_hi);
        }
        if (rem >= 3) {
            const int e = e_lo + it + 2;
            fetch_g(2, true, g2);
            sample_checked(e < e_hi ? lds4(ent_s + e * 16) : z4, g2, e < e_hi);
        }
        if (DEFER && SPLIT >= 4) {
            finish(d_a, d_h, d_w, d_idx);
            d_idx = -1;
        }
        if (GRING) asm volatile("cp.async.wait_group 0;" ::: "memory");  // nothing may land in a slot during the next unit
        if (cur >= 0) flush();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// (the A/B switches of round 1 -- lane splits 2 / 4, the cp.async grad_out ring, the deferred reduction, L2 prefetch -- all lost
// against this configuration and are no longer instantiated: DESIGN.md section 5.6 keeps their measurements)
constexpr int kSplit = 8, kModeLean = 16, kModeLeanFused = 20;

bool make_params(const Problem &p, SortParams &sp, size_t &smem)
{
    if (!((p.dtype == 0 || p.dtype == 1) && p.D == kD && p.have_host && p.L <= kMaxLevels)) return false;
    if (p.Lq < 1 || p.Lq > 65535 || p.P > 16 || (long long)p.Lq * p.P > kMaxSamples) return false;
    sp.L = p.L; sp.P = p.P; sp.S = p.S; sp.M = p.M; sp.Lq = p.Lq;
    for (int l = 0; l < p.L; ++l) {
        sp.H[l] = p.table.H[l];
        sp.W[l] = p.table.W[l];
        sp.start[l] = p.table.start[l];
        int cs = 0;
        while ((1 << cs) < sp.W[l] + 1) ++cs;
        sp.cshift[l] = cs;
        if (((long long)(sp.H[l] + 1) << cs) > kMaxCells) return false;
        if ((long long)(sp.H[l] + 2) * (sp.W[l] + 2) * p.M * kD >= (1LL << 31)) return false;  // 32-bit row offsets
    }
    sp.vrows_max = 0;
    for (int l = 0; l < p.L; ++l) {
        const int rows = (sp.H[l] + 2) * (sp.W[l] + 2);
        if (rows > sp.vrows_max) sp.vrows_max = rows;
        if ((long long)sp.H[l] * sp.W[l] >= 65536) return false;
        sp.wrecip[l] = (unsigned)((0x100000000ULL + sp.W[l] - 1) / sp.W[l]);
    }
    sp.units = (long long)p.N * p.M * p.L;
    if (sp.units >= (1LL << 30)) return false;  // 32-bit unit arithmetic in the kernel
    smem = (size_t)sp.vrows_max * 8 * 16 + ((size_t)p.Lq * p.P + 4) * 16 + (kMaxCells + 8 + 32) * sizeof(int) + kMaxCells * sizeof(int2);
    if (smem > (size_t)p.smem_optin) return false;
    sp.prefetch = 0;
    sp.fused = p.ref != nullptr ? 1 : 0;
    if (sp.fused && p.loc_dtype != 0) return false;
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

// schedule: `grid` persistent CTAs take `base` whole units each; the `rem` leftover units are cut into `parts` pieces, one per CTA
void plan_units(const Problem &p, SortParams &sp, int &grid)
{
    grid = (int)(sp.units < p.sm_count ? sp.units : p.sm_count);
    sp.base = (int)(sp.units / grid);
    sp.rem = (int)(sp.units - (long long)sp.base * grid);
    sp.parts = sp.rem > 0 ? grid / sp.rem : 1;
    if (sp.parts > 4) sp.parts = 4;
    if (sp.parts < 1) sp.parts = 1;
    // whole units zero their own grad_value slab in the kernel; the leftover units start at unit grid*base, i.e. inside frame
    // (grid*base / L) / M: that frame and everything after it is zeroed by the host instead
    sp.zero_frames = sp.rem > 0 ? (int)(((long long)grid * sp.base / p.L) / p.M) : p.N;
}

template <typename TV, typename TM>
cudaError_t launch_sorted(const Problem &p, const BackwardIO &io, float *gv_acc, SortParams &sp, size_t smem, int grid, cudaStream_t st)
{
    auto go = [&](auto kern_c) {
        constexpr auto kern = decltype(kern_c)::value;
        cudaError_t e = ensure_smem_optin<kern>(p);
        if (e != cudaSuccess) return e;
        kern<<<grid, threads_for_split(kSplit), smem, st>>>((const TV *)p.value, (const TM *)p.loc, (const TM *)p.attn, (const TV *)io.grad_out, gv_acc,
                                                          (TM *)io.grad_loc, (TM *)io.grad_attn, (const TM *)p.ref, p.mask, sp);
        count_launch();
        return cudaGetLastError();
    };
    using std::integral_constant;
#define MSDA_SORTED_KERN(PT, MD) integral_constant<decltype(&bwd_sorted32_kernel<TV, TM, PT, kSplit, MD>), &bwd_sorted32_kernel<TV, TM, PT, kSplit, MD>>()
    if constexpr (std::is_same<TV, float>::value && std::is_same<TM, float>::value) {  // the fused prologue is float32-only
        if (sp.fused) return p.P == 4 ? go(MSDA_SORTED_KERN(4, kModeLeanFused)) : go(MSDA_SORTED_KERN(0, kModeLeanFused));
    }
    if (sp.fused) return cudaErrorInvalidValue;
    return p.P == 4 ? go(MSDA_SORTED_KERN(4, kModeLean)) : go(MSDA_SORTED_KERN(0, kModeLean));
#undef MSDA_SORTED_KERN
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
    float *acc = p.dtype == 0 ? (float *)io.grad_value : (float *)io.workspace;  // bf16: fp32 accumulation in the workspace
    int grid = 1;
    if (sp.units > 0 && p.Lq > 0) plan_units(p, sp, grid);
    else sp.zero_frames = 0;
    // the kernel zeroes what whole units own -- provided the levels tile [0, S) exactly, so that every pixel belongs to a unit
    long long next = 0;
    for (int l = 0; l < p.L; ++l) {
        if (sp.start[l] != next) sp.zero_frames = 0;
        next += (long long)sp.H[l] * sp.W[l];
    }
    if (next != p.S) sp.zero_frames = 0;
    if (n_val > 0 && sp.zero_frames < p.N) {
        const size_t frame = (size_t)p.S * p.M * p.D;
        cudaError_t e = cudaMemsetAsync(acc + (size_t)sp.zero_frames * frame, 0, sizeof(float) * frame * (size_t)(p.N - sp.zero_frames), st);
        if (e != cudaSuccess) return e;
    }
    if (sp.units == 0 || p.Lq == 0) return cudaSuccess;
    cudaError_t e = p.dtype == 0      ? launch_sorted<float, float>(p, io, acc, sp, smem, grid, st)
                    : p.loc_dtype == 0 ? launch_sorted<bf16, float>(p, io, acc, sp, smem, grid, st)
                                       : launch_sorted<bf16, bf16>(p, io, acc, sp, smem, grid, st);
    if (e != cudaSuccess) return e;
    if (p.dtype == 1) return convert_f32_to_bf16(acc, io.grad_value, n_val, p.sm_count, st);
    return cudaSuccess;
}

}  // namespace msda
