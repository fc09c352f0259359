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
//   - per sample only ONE 128-byte row is read: its grad_output row, streamed from L2 and prefetched two samples ahead
//     into rotating register sets,
//   - grad_sampling_loc / grad_attn_weight are 8-lane shuffle reductions (transposed butterfly, 4 shuffles) and plain stores.
// The zero-padded value slab of the level sits in shared memory, so a cell switch costs 4 LDS.128, not 4 L2 round trips.
// MODE bit 3 (float32 values): the slab is moved by the TMA unit -- ONE cp.async.bulk.tensor issued by one thread per unit, a
// (H+2) x (W+2) pixel box at (-1, -1) whose out-of-bounds border arrives as zeros, completion on an mbarrier -- instead of
// eight cp.async + index arithmetic per thread and a hand-zeroed border (profiles/r06a: that staging phase held 16.6 % of the
// kernel's stall samples, throttled on the load/store queue).
// A lane group is 8 lanes x 4 channels (one 128-byte row per LDS.128 phase: conflict-free).
// (The A/B switches of round 1 -- lane splits 2 / 4, a cp.async grad_out ring, a deferred reduction, L2 prefetch of the next
// unit -- all lost against this configuration and are gone: DESIGN.md section 5.6 keeps their measurements.)
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "msda_kernels.h"
#include "msda_tma.cuh"

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

// A sample is shared by 8 lanes; each lane owns one chunk of 4 channels.
// TV: value / grad_out element type (float or bf16; grad_value always accumulates in fp32: for bf16 into a workspace that a
// second kernel narrows).  TM: element type of sampling locations, attention weights and their gradients.
// MODE: kModeFused (4) = module prologue folded in; kModeTma (8) = the slab is staged by the TMA unit (float32 values only).
constexpr int kModeFused = 4, kModeTma = 8;
constexpr int kSplit = 8, kThreads = 768;

// The walk of one lane over its group's entry range [e_lo, e_hi) of the sorted list.  Everything the loop needs arrives in one
// struct, so that what the kernel around it keeps alive (unit decode, TMA issue, sort passes) is visibly separate from the loop's
// own state: ptxas is capped at 80 registers by the 768-thread CTA, and values that leak into the loop cost issue slots there.
struct WalkArgs {
    unsigned ent_s, cinfo_s, V_s;  // shared-memory byte addresses: sorted entries, per-cell table, value slab (+ this lane's chunk)
    int Wpb;                       // bytes per row of the padded slab
    int e_lo, e_hi, full;          // this group's entries; `full` = entries every group owns (the others own one more)
    const void *gobase;            // grad_out row 0 of this (frame, head): the same for every lane (a uniform register pair)
    unsigned grow_b;               // bytes between grad_out rows of consecutive queries
    unsigned glane_b;              // byte offset of this lane's chunk inside a row
    float *gbase;                  // grad_value pixel (0, 0) of this (frame, head, level), this lane's chunk
    int o01, o10;                  // grad_value element offsets of the right / lower neighbour
    void *res_base;                // where this lane's result of a sample goes (grad_attn, grad_loc.y or grad_loc.x; lanes 3..7: unused)
    int res_stride;
    float res_scale;
    int qs, j;                     // samples between consecutive queries of one head; lane index inside the group
};

template <typename TV, typename TM>
__device__ __forceinline__ void walk_range(const WalkArgs &wa)
{
    const unsigned ent_s = wa.ent_s, cinfo_s = wa.cinfo_s, V_s = wa.V_s;
    const int Wpb = wa.Wpb, e_lo = wa.e_lo, e_hi = wa.e_hi, qs = wa.qs, j = wa.j;
    const TV *gobase = reinterpret_cast<const TV *>(wa.gobase);
    const unsigned grow_b = wa.grow_b, glane_b = wa.glane_b;
    float *gbase = wa.gbase;
    const int o01 = wa.o01, o10 = wa.o10, o11 = wa.o01 + wa.o10;
    TM *res_base = reinterpret_cast<TM *>(wa.res_base);
    const int res_stride = wa.res_stride;
    const float res_scale = wa.res_scale;
    int cur = -1, inmask = 0;
    float *gp = gbase;             // grad_value row (this lane's chunk) of the current cell's top-left neighbour
    f4p v00, dW, dH, dd;           // V00, V01-V00, V10-V00, V11-V10-V01+V00 of the current cell
    // grad_value of the current cell in the basis (1, lw, lh, lh*lw): sums of att * basis * grad_out
    const float2 z2 = make_float2(0.f, 0.f);
    f4p a00 = {z2, z2}, a01 = {z2, z2}, a10 = {z2, z2}, a11 = {z2, z2};

    // (S0, S1, S2, S3) -> the four bilinear weights, one vector reduction per neighbour row inside the map
    auto flush = [&]() {
        const f4p r10 = psub(a10, a11);                    // lh*(1-lw)
        const f4p r00 = psub(psub(a00, a01), r10);         // (1-lh)*(1-lw) = S0 - S1 - (S2 - S3)
        const f4p r01 = psub(a01, a11);                    // (1-lh)*lw
        if (inmask == 15) {  // interior cell (86 % of a 28 x 28 level): no predicates, no reconvergence points
            red_add_v4(gp, r00.lo.x, r00.lo.y, r00.hi.x, r00.hi.y);
            red_add_v4(gp + o01, r01.lo.x, r01.lo.y, r01.hi.x, r01.hi.y);
            red_add_v4(gp + o10, r10.lo.x, r10.lo.y, r10.hi.x, r10.hi.y);
            red_add_v4(gp + o11, a11.lo.x, a11.lo.y, a11.hi.x, a11.hi.y);
        } else {
            red_add_v4_if(inmask & 1, gp, r00.lo.x, r00.lo.y, r00.hi.x, r00.hi.y);
            red_add_v4_if(inmask & 2, gp + o01, r01.lo.x, r01.lo.y, r01.hi.x, r01.hi.y);
            red_add_v4_if(inmask & 4, gp + o10, r10.lo.x, r10.lo.y, r10.hi.x, r10.hi.y);
            red_add_v4_if(inmask & 8, gp + o11, a11.lo.x, a11.lo.y, a11.hi.x, a11.hi.y);
        }
    };
    // start the load of the grad_out chunk of sorted entry e2 into the register set g
    auto issue_g = [&](int e2, float4 &g) {
        unsigned packed;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(packed) : "r"(ent_s + e2 * 16 + 12));
        // 64-bit base + unsigned 32-bit byte offset: two adds instead of a sign-extended 64-bit multiply-shift chain
        // uniform 64-bit base + 32-bit (row * pitch + lane chunk): one multiply-add and one 64-bit add per load
        const TV *gq = reinterpret_cast<const TV *>(reinterpret_cast<const char *>(gobase) + ((packed & 0xffff) * grow_b + glane_b));
        g = load4(gq);
    };
    // s_h / s_w arrive already multiplied by the sample's attention weight (cuh:126-127), so the tail is one multiply per lane
    // lane-dependent selects of the butterfly as ONE bitwise instruction each on a mask kept in a register: as predicates, j & 1 and
    // j & 2 were re-derived in every step (three instructions; there are only seven predicate registers)
    unsigned m1 = (j & 1) ? 0xffffffffu : 0u, m2 = (j & 2) ? 0xffffffffu : 0u;
    pin(m1);
    pin(m2);
    auto pick = [](unsigned m, float if_set, float if_clear) {
        unsigned r;
        asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(m), "r"(__float_as_uint(if_set)), "r"(__float_as_uint(if_clear)));
        return __uint_as_float(r);
    };
    auto finish = [&](float s_a, float s_h, float s_w, int sidx) {  // transposed butterfly + stores
        float x = pick(m1, s_h, s_a);
        x += __shfl_xor_sync(0xffffffffu, pick(m1, s_a, s_h), 1);
        s_w += __shfl_xor_sync(0xffffffffu, s_w, 1);
        float keep = pick(m2, s_w, x);
        keep += __shfl_xor_sync(0xffffffffu, pick(m2, x, s_w), 2);
        keep += __shfl_xor_sync(0xffffffffu, keep, 4);
        if (sidx >= 0 && j < 3) store1(res_base + sidx * res_stride, keep * res_scale);
    };
    // One sample: en = its sorted entry, gr = its grad_out chunk (both already in registers).  Fewest FMA-pipe instructions per
    // sample: four dot products of grad_out with (V00, dW, dH, dd) combined after the horizontal add; grad_value accumulated in the
    // basis (1, lw, lh, lh*lw) and converted to the four neighbours once per cell at flush.  The first sample of a cell
    // STARTS the accumulators (no zeroing at the switch: 22 instructions per switch in the previous SASS).
#define MSDA_SAMPLE_BODY \
        float s_a = 0.f, s_h = 0.f, s_w = 0.f, att_v = 0.f; \
        int sidx = 0; \
        if (active) { \
            const int packed = __float_as_int(en.w); \
            const int cell = packed >> 20; \
            sidx = (packed & 0xffff) * qs + ((packed >> 16) & 0xf); \
            att_v = en.z; \
            const float lh = en.x, lw = en.y; \
            const float a1s = att_v * lw, a2s = att_v * lh, a3s = a2s * lw; \
            const float2 b0 = make_float2(att_v, att_v), b1 = make_float2(a1s, a1s), b2 = make_float2(a2s, a2s), b3 = make_float2(a3s, a3s); \
            const f4p g = pk(gr); \
            const bool cell_same = cell == cur; \
            if (!cell_same) { \
                flush(); \
                cur = cell; \
                int2 ci; \
                asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(ci.x), "=r"(ci.y) : "r"(cinfo_s + cell * 8)); \
                inmask = ci.x & 15; \
                gp = gbase + ci.y; \
                const unsigned a = V_s + (unsigned)(ci.x & ~127); \
                const f4p t00 = pk(lds4(a)), t01 = pk(lds4(a + kD * 4)), t10 = pk(lds4(a + Wpb)), t11 = pk(lds4(a + Wpb + kD * 4)); \
                v00 = t00; \
                dW = psub(t01, t00); \
                dH = psub(t10, t00); \
                dd = psub(psub(t11, t10), dW); \
                a00.lo = __fmul2_rn(b0, g.lo); a00.hi = __fmul2_rn(b0, g.hi); \
                a01.lo = __fmul2_rn(b1, g.lo); a01.hi = __fmul2_rn(b1, g.hi); \
                a10.lo = __fmul2_rn(b2, g.lo); a10.hi = __fmul2_rn(b2, g.hi); \
                a11.lo = __fmul2_rn(b3, g.lo); a11.hi = __fmul2_rn(b3, g.hi); \
            } else { \
                a00.lo = __ffma2_rn(b0, g.lo, a00.lo); a00.hi = __ffma2_rn(b0, g.hi, a00.hi); \
                a01.lo = __ffma2_rn(b1, g.lo, a01.lo); a01.hi = __ffma2_rn(b1, g.hi, a01.hi); \
                a10.lo = __ffma2_rn(b2, g.lo, a10.lo); a10.hi = __ffma2_rn(b2, g.hi, a10.hi); \
                a11.lo = __ffma2_rn(b3, g.lo, a11.lo); a11.hi = __ffma2_rn(b3, g.hi, a11.hi); \
            } \
            const float2 qA = __ffma2_rn(g.hi, v00.hi, __fmul2_rn(g.lo, v00.lo)); \
            const float2 qB = __ffma2_rn(g.hi, dW.hi, __fmul2_rn(g.lo, dW.lo)); \
            const float2 qC = __ffma2_rn(g.hi, dH.hi, __fmul2_rn(g.lo, dH.lo)); \
            const float2 qE = __ffma2_rn(g.hi, dd.hi, __fmul2_rn(g.lo, dd.lo)); \
            const float A = qA.x + qA.y, B = qB.x + qB.y, C = qC.x + qC.y, E = qE.x + qE.y; \
            s_w = __fmaf_rn(lh, E, B); \
            s_h = __fmaf_rn(lw, E, C); \
            s_a = __fmaf_rn(lh, s_h, __fmaf_rn(lw, B, A)); \
        } \
        finish(s_a, s_h * att_v, s_w * att_v, active ? sidx : -1);
    // nvcc 12.9 crashes on a generic lambda here, hence two plain lambdas sharing one body
    auto sample_steady = [&](const float4 &en, const float4 &gr) {
        const bool active = true;  // steady-state steps are active for every group
        MSDA_SAMPLE_BODY
    };
    auto sample_checked = [&](const float4 &en, const float4 &gr, bool active) {
        MSDA_SAMPLE_BODY
    };
#undef MSDA_SAMPLE_BODY

    // software pipeline: grad_out rows are loaded TWO samples ahead into three rotating register sets (the loop is unrolled
    // by three so the rotation costs no moves); one sample ahead does not cover the L2 latency.  Every group owns `full`
    // or `full`+1 entries, so the first `full` steps need no activity checks; the remainder is peeled.
    const int full = wa.full, cnt = e_hi - e_lo;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 g0 = z4, g1 = z4, g2 = z4;
    if (cnt > 0) {
        issue_g(e_lo, g0);
        issue_g(e_lo + 1, g1);
    }
    int it = 0;
    for (; it + 3 <= full; it += 3) {
        const int e = e_lo + it;
        issue_g(e + 2, g2);
        sample_steady(lds4(ent_s + e * 16), g0);
        issue_g(e + 3, g0);
        sample_steady(lds4(ent_s + (e + 1) * 16), g1);
        issue_g(e + 4, g1);
        sample_steady(lds4(ent_s + (e + 2) * 16), g2);
    }
    // 0..3 remaining steps, activity checked.  (The count is the same for the four groups of a warp up to one step; all of them take
    // part in the shuffles, so the warp runs the largest count.)
    const int rem = __reduce_max_sync(0xffffffffu, cnt - it);
    if (rem >= 1) {
        const int e = e_lo + it;
        if (cnt > it + 2) issue_g(e + 2, g2);
        sample_checked(e < e_hi ? lds4(ent_s + e * 16) : z4, g0, e < e_hi);
    }
    if (rem >= 2) {
        const int e = e_lo + it + 1;
        sample_checked(e < e_hi ? lds4(ent_s + e * 16) : z4, g1, e < e_hi);
    }
    if (rem >= 3) {
        const int e = e_lo + it + 2;
        sample_checked(e < e_hi ? lds4(ent_s + e * 16) : z4, g2, e < e_hi);
    }
    if (cur >= 0) flush();
}

template <typename TV, typename TM, int P_T, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
bwd_sorted32_kernel(const TV *__restrict__ value, const TM *__restrict__ loc, const TM *__restrict__ attn,
                    const TV *__restrict__ grad_out, float *__restrict__ grad_value, TM *__restrict__ grad_loc,
                    TM *__restrict__ grad_attn, const TM *__restrict__ ref, const unsigned char *__restrict__ mask,
                    const __grid_constant__ SortParams sp, const __grid_constant__ TmaMaps maps)
{
    constexpr int SPLIT = kSplit;
    constexpr int kGroups = kThreads / SPLIT, kRounds = (kMaxSamples + kThreads - 1) / kThreads;
    constexpr int kCPT = (kMaxCells + kThreads - 1) / kThreads;  // cells scanned per thread
    constexpr bool FUSED = (MODE & kModeFused) != 0, TMA = (MODE & kModeTma) != 0 && sizeof(TV) == 4;
    extern __shared__ __align__(128) float4 smem4[];
    const int Lq = sp.Lq, P = P_T > 0 ? P_T : sp.P, EP = Lq * P, LP = sp.L * P;
    float4 *V4 = smem4;                                   // zero-padded value slab of the current level: (H+2) x (W+2) rows
    float4 *ent = smem4 + (size_t)sp.vrows_max * 8;       // sorted samples: {lh, lw, attn, cell<<20 | p<<16 | q}
    int *cell_off = reinterpret_cast<int *>(ent + EP + 4);  // (four spare entries, see the prefetch) kMaxCells + 1 counters -> exclusive offsets
    int *warp_tot = cell_off + kMaxCells + 8;
    // per-cell facts of the current level, filled while the unit is sorted: {value-slab byte offset of the cell's top-left
    // neighbour | 4-bit mask of the neighbours inside the map, grad_value element offset of that neighbour}
    int2 *cinfo = reinterpret_cast<int2 *>(warp_tot + 32);
    unsigned cinfo_s = (unsigned)__cvta_generic_to_shared(cinfo);
    pin(cinfo_s);
    const unsigned vbar = (unsigned)__cvta_generic_to_shared(cinfo + kMaxCells);  // mbarrier: the slab of the current unit has landed
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int grp = tid / SPLIT;
    int j = tid % SPLIT;                                  // lane j of its sample group
    pin(j);
    int chf = j * 4;                                      // float offset of this lane's chunk inside a 32-channel row
    pin(chf);
    unsigned ent_s = (unsigned)__cvta_generic_to_shared(ent);
    pin(ent_s);
    const int pix = sp.M * kD;
    const int qstride = sp.M * LP;  // samples between consecutive queries of one head
    unsigned vphase = 0;
    const int n_mine = sp.base + ((int)blockIdx.x < sp.rem * sp.parts ? 1 : 0);
    if (TMA && tid == 0) {
        mbar_init(vbar, 1);
        mbar_fence_init();
        for (int l = 0; l < sp.L; ++l) tma_prefetch_map(&maps.level[l]);
    }

    // schedule: in round ui, CTA b runs whole unit ui*grid + b; then, if it exists, piece b of the leftover units (leftover unit
    // grid*base + b/parts, entry range part b%parts of `parts`): without the split the CTAs that get one unit more than the others
    // set the kernel time (960 units over 148 CTAs: 7 vs 6.49).
    // The rounds are INTERLEAVED over the CTAs (not a contiguous range per CTA) so that at any time the grid works on ~grid/L
    // consecutive (frame, head) pairs: the three levels of a pair run concurrently on three SMs and share its grad_out rows, its
    // location / weight lines and the lines of grad_loc / grad_attn through L2 while they are hot.  With contiguous ranges every CTA
    // had its own pair in flight -- 2 x 148 x 0.7 MB of reusable lines against a 126 MB L2 that is also streaming 220 MB -- and
    // DRAM moved 342 MB per launch for 221 MB of algorithmic traffic; interleaved: 179 MB, and 198 -> 178 us (profiles/r08b).
    for (int ui = 0; ui < n_mine; ++ui) {
        const bool piece = ui == sp.base;
        // (32-bit unit arithmetic: every thread decodes its unit, and 64-bit divisions are ~100 instructions each)
        const int u = piece ? (int)gridDim.x * sp.base + (int)blockIdx.x / sp.parts : ui * (int)gridDim.x + (int)blockIdx.x;
        const int part = piece ? (int)blockIdx.x % sp.parts : 0, parts = piece ? sp.parts : 1;
        const int nm = u / sp.L;
        const int l = u - nm * sp.L;
        const long long n = nm / sp.M;
        const int m = nm - (int)n * sp.M;
        const int H = sp.H[l], W = sp.W[l], cs = sp.cshift[l], cells = (H + 1) << cs, Wp = W + 2;
        const float Hf = (float)H, Wf = (float)W;
        const long long lvl_row = (n * sp.S + sp.start[l]) * (long long)pix + (long long)m * kD;  // pixel (0,0), channel 0
        const TV *vbase = value + lvl_row;
        // What does not touch shared memory goes BEFORE the barrier that ends the previous unit: the loads of this unit's locations and
        // weights (into registers) and the zeroing of its grad_value slab are issued by every warp as soon as IT has finished its part of
        // the previous walk, so their latency runs under the wait for the slowest group instead of after it (that barrier alone held
        // 4.5 % of the kernel's stall samples, the location loads' long-scoreboard wait another 4.5 %: profiles/r08c_*before_reorder.txt and its source page): 177.7 -> 171.7 us.
        // (Going further -- the per-cell table and the whole counting pass ahead of the barrier too, into a second set of counters --
        // LOSES: 184 us; the early warps' shared-memory atomics and stores compete with the warps that are still walking, and those
        // set the time.  profiles/r08a-r_*, profiles/experiments/r08f_*.patch)
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
        // A unit this CTA owns alone zeroes its own slab of grad_value here, right before it accumulates into it: the lines are
        // still in L2 when the reductions arrive, and no 42 MB memset runs ahead of the kernel.  (The few leftover units that
        // several CTAs share lie in the frames >= zero_frames, which the host zeroes with a small memset.)
        if (!piece && n < sp.zero_frames) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            float *gvl = grad_value + lvl_row;
            for (int i = tid; i < H * W * 8; i += kThreads) *reinterpret_cast<float4 *>(gvl + (i >> 3) * pix + (i & 7) * 4) = z;
        }
        if (TMA && FUSED && mask != nullptr) fence_proxy_async();  // the masked rows zeroed below (generic proxy) precede the next TMA fill
        __syncthreads();  // previous unit is completely finished with V / ent / cell_off
        if (TMA && tid == 0) {  // one instruction moves the level; the border of the box lies outside the tensor and arrives as zeros
            mbar_arrive_expect_tx(vbar, (unsigned)((H + 2) * Wp) * (kD * 4));
            // (L2 evict-first: the level is read once, while grad_out rows, locations and the grad_value lines under reduction are
            // re-used from L2 for as long as the (frame, head) is worked on -- 200.6 -> 194.8 us, profiles/r08a)
            tma_load_5d_hint((unsigned)__cvta_generic_to_shared(V4), &maps.level[l], vbar, 0, m, -1, -1, (int)n, l2_policy_evict_first());
        }
        // (the slab copy and the zero fills below overlap the latency of the loads just issued)
        if (!TMA) {
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
        if (TMA) {
            mbar_wait(vbar, vphase);
            vphase ^= 1;
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (FUSED && mask != nullptr) {  // padding mask folded in: masked pixels read as zero rows (their grad_value is zeroed afterwards)
            if (!TMA) __syncthreads();
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

        // ---- walk the sorted list: every 8-lane group owns a contiguous range of entries ------------------------------------
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
        // every 8-lane group owns a contiguous range of entries; `full` or `full`+1 each
        WalkArgs wa;
        wa.ent_s = ent_s;
        wa.cinfo_s = cinfo_s;
        wa.V_s = (unsigned)__cvta_generic_to_shared(V4) + chf * 4;
        wa.Wpb = Wp * (kD * 4);
        // Which range a group walks: the r-th of kGroups equal ranges with r = (group within its warp) * (warps) + warp, NOT r = group.
        // The first and the last row of cells of a level only receive samples from half their area -- twice the cell switches per
        // sample -- and a warp pays for a switch whenever ANY of its four groups switches: with r = group, warp 0 (ranges 0..3, all in
        // the first rows) was the last to finish in 52 % of the units and the last warp in 18 %, the others waiting ~10 % of the walk
        // at the unit barrier (globaltimer stamps per warp, profiles/r08a-r_*).  Transposed, every warp holds at most one such range:
        // 172.1 -> 170.2 us (N=80: 322.4 -> 318.6).  (Steering those ranges to the high warp ids that the arbiter favours: no further gain.)
        const int rng = (grp & 3) * (kGroups / 4) + (grp >> 2);
        wa.e_lo = E0 + E * rng / kGroups;  // E <= 4608, rng < 96: fits 32 bits
        wa.e_hi = E0 + E * (rng + 1) / kGroups;
        wa.full = E / kGroups;
        wa.gobase = grad_out + (n * Lq * sp.M + m) * (long long)kD;       // grad_out row of query q: gobase + q*pix
        wa.grow_b = (unsigned)pix * (unsigned)sizeof(TV);                  // bytes between the grad_out rows of consecutive queries
        wa.glane_b = (unsigned)chf * (unsigned)sizeof(TV);
        wa.gbase = grad_value + lvl_row + chf;
        wa.o01 = pix;
        wa.o10 = W * pix;
        // result stores: lane 0 grad_attn, lane 1 grad_loc.y, lane 2 grad_loc.x
        wa.res_base = j == 0 ? (void *)gatt : (void *)(gloc + (j == 1 ? 1 : 0));
        wa.res_stride = j == 0 ? 1 : 2;
        // d(loc)/d(offset) = 1 / (W, H) cancels the (W, H) factor of grad_sampling_loc (cuh:126-127) in the fused form
        wa.res_scale = (j == 0 || FUSED) ? 1.f : (j == 1 ? Hf : Wf);
        wa.qs = qstride;
        wa.j = j;
        walk_range<TV, TM>(wa);
    }
    if (!TMA) asm volatile("cp.async.wait_group 0;" ::: "memory");
}

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
    smem = (size_t)sp.vrows_max * 8 * 16 + ((size_t)p.Lq * p.P + 4) * 16 + (kMaxCells + 8 + 32) * sizeof(int) + kMaxCells * sizeof(int2) + 16;  // (+ the slab mbarrier)
    if (smem > (size_t)p.smem_optin) return false;
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

// schedule: `grid` persistent CTAs take `base` whole units each (round-robin: unit ui*grid + b); the `rem` leftover units are cut
// into `parts` pieces, one per CTA
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

// Tensor maps of the levels for the TMA staging; false when a level does not fit a box, the pointer is misaligned, the driver
// lacks cuTensorMapEncodeTiled or MSDA_SORTED_TMA=0 asks for the cp.async staging (A/B switch).
bool make_level_maps(const Problem &p, const SortParams &sp, TmaMaps &maps)
{
    static const bool off = [] { const char *e = getenv("MSDA_SORTED_TMA"); return e && atoi(e) == 0; }();
    if (off || p.dtype != 0 || p.L > kTmaMaxLevels) return false;
    for (int l = 0; l < p.L; ++l)
        if (!tma_make_level_map(&maps.level[l], (const float *)p.value + (size_t)sp.start[l] * p.M * kD, p.N, p.S, p.M, sp.H[l], sp.W[l]))
            return false;
    return true;
}

template <typename TV, typename TM>
cudaError_t launch_sorted(const Problem &p, const BackwardIO &io, float *gv_acc, SortParams &sp, size_t smem, int grid, cudaStream_t st)
{
    TmaMaps maps;
    bool tma = false;
    if constexpr (std::is_same<TV, float>::value) tma = make_level_maps(p, sp, maps);
    auto go = [&](auto kern_c) {
        constexpr auto kern = decltype(kern_c)::value;
        cudaError_t e = ensure_smem_optin<kern>(p);
        if (e != cudaSuccess) return e;
        kern<<<grid, kThreads, smem, st>>>((const TV *)p.value, (const TM *)p.loc, (const TM *)p.attn, (const TV *)io.grad_out, gv_acc,
                                           (TM *)io.grad_loc, (TM *)io.grad_attn, (const TM *)p.ref, p.mask, sp, maps);
        count_launch();
        return cudaGetLastError();
    };
    using std::integral_constant;
#define MSDA_SORTED_KERN(PT, MD) integral_constant<decltype(&bwd_sorted32_kernel<TV, TM, PT, MD>), &bwd_sorted32_kernel<TV, TM, PT, MD>>()
    if constexpr (std::is_same<TV, float>::value && std::is_same<TM, float>::value) {  // the fused prologue and the TMA staging are float32-only
        if (sp.fused) {
            if (tma) return p.P == 4 ? go(MSDA_SORTED_KERN(4, kModeFused | kModeTma)) : go(MSDA_SORTED_KERN(0, kModeFused | kModeTma));
            return p.P == 4 ? go(MSDA_SORTED_KERN(4, kModeFused)) : go(MSDA_SORTED_KERN(0, kModeFused));
        }
        if (tma) return p.P == 4 ? go(MSDA_SORTED_KERN(4, kModeTma)) : go(MSDA_SORTED_KERN(0, kModeTma));
    }
    if (sp.fused) return cudaErrorInvalidValue;
    return p.P == 4 ? go(MSDA_SORTED_KERN(4, 0)) : go(MSDA_SORTED_KERN(0, 0));
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
