// msda_pipe32.cu -- D == 32, float32, geometry known on the host: sort-and-segment backward, SOFTWARE-PIPELINED.
//
// Same mathematics and the same owner-walks-a-sorted-list design as msda_sorted32.cu (reference: ms_deform_im2col_cuda.cuh
// :92-164 bilinear backward, :306-408 kernel).  What changes is the schedule, for two reasons ncu gave on sorted32
// (profiles/r01p_*, profiles/r01q_*):
//  (1) the counting sort of a (frame, head, level) unit -- load locations, histogram, scan, fill, slab copy -- is five
//      barrier-separated, latency-bound phases that keep all 24 warps waiting for 28 % of the kernel time while issuing
//      17 % of its instructions.  Here the CTA is warp-specialised: SORTER warps prepare unit u+1 (samples staged in shared
//      memory with one cp.async round trip, histogram with shared-memory integer atomics, scan, fill) while WALKER warps
//      walk the sorted list of unit u; buffers are handed over with named barriers (bar.arrive / bar.sync, ids 1-4), so a
//      walker never sits in a __syncthreads.
//  (2) ptxas puts every grad_out prefetch LDG on ONE scoreboard, so the first use of any of them waits for all of them,
//      including the load issued a few instructions earlier: the software prefetch never ran ahead (20 % long-scoreboard
//      stalls).  Here every lane copies its 16-byte chunk of the grad_out row of the sample three steps ahead into a
//      private shared-memory slot with cp.async and reads it back with LDS.128 after cp.async.wait_group 3 -- a counted,
//      in-order wait that lets the three newer copies stay in flight.
// The sorted entries {lh, lw, attn, cell|p|q} (16 B) live in a per-CTA, double-buffered global scratch that stays in L2
// (two units of them do not fit in shared memory beside the value slabs); every group streams its own contiguous range
// through a two-block shared-memory ring (one coalesced 128-byte block per eight samples, riding in the same cp.async
// groups).  The per-sample reduction + store of sample i is deferred into the basic block of sample i+1, where it
// overlaps that sample's FMAs.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kD = 32;
constexpr int kMaxSamples = 4608;   // Lq*P per level (bounds the scratch and keeps q in 16 bits)
constexpr int kMaxCells = 1024;     // cell-grid size per level
constexpr int kThreads = 768;       // 24 warps: SW sorter warps + (24 - SW) walker warps
constexpr int kBlk = 8;             // sorted entries per ring block
constexpr int kAhead = 3;           // grad_out rows are copied this many steps ahead (kAhead + 1 slots per lane)
constexpr int kMaxCtas = 192;       // scratch is sized for this many CTAs (B200: 148 SMs)
constexpr int kScratchPad = 32;     // entries a ring block may read past the end of a list
enum { kBarFull0 = 1, kBarFull1 = 2, kBarEmpty0 = 3, kBarEmpty1 = 4, kBarSort = 5 };

struct PipeParams {
    int L, P, S, M, Lq;
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    int cshift[kMaxLevels];       // log2 of the cell-grid pitch (power of two >= W+1): cell = (y0+1) << cshift | (x0+1)
    unsigned wrecip[kMaxLevels];  // ceil(2^32 / W)
    int vrows[kMaxLevels];        // rows of the zero-padded level slab, (H+2)*(W+2)
    int vrows_total;              // rows of the ping-pong value region: unit i sits at its low end (i even) or high end (i odd)
    int stage_q;                  // queries whose samples the sorter stages at a time (shared-memory budget)
    int scratch_stride;           // entries per scratch buffer
    long long units;              // N*M*L
    int base, rem, parts;         // schedule, as in sorted32
};

__device__ __forceinline__ void cp_async16(unsigned smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

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
__device__ __forceinline__ void pin(float &x) { asm volatile("" : "+f"(x)); }
template <typename T> __device__ __forceinline__ void pin(T *&x) { asm volatile("" : "+l"(x)); }
__device__ __forceinline__ float4 lds4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds1(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

struct Unit {
    long long n;
    int m, l, part, parts;
};

// SW = 2: 2 sorter + 22 walker warps, 80 registers each.  SW = 4: one sorter warpgroup that gives registers back
// (setmaxnreg.dec 40: 128 x 40 registers freed = 640 x 8 taken) so that the five walker warpgroups can take 88 (setmaxnreg.inc).
template <int P_T, int SPLIT, bool DEFER, int SW>
__global__ void __launch_bounds__(kThreads, 1)
bwd_pipe32_kernel(const float *__restrict__ value, const float *__restrict__ loc, const float *__restrict__ attn,
                  const float *__restrict__ grad_out, float *__restrict__ grad_value, float *__restrict__ grad_loc,
                  float *__restrict__ grad_attn, float4 *__restrict__ scratch, const __grid_constant__ PipeParams sp)
{
    constexpr int CH = 8 / SPLIT;
    constexpr int kSortWarps = SW, kSortThreads = SW * 32, kWalkThreads = kThreads - kSortThreads;
    constexpr int kGroups = kWalkThreads / SPLIT;
    extern __shared__ float4 smem4[];
    const int Lq = sp.Lq, P = P_T > 0 ? P_T : sp.P, LP = sp.L * P;
    float4 *V4 = smem4;                                                  // ping-pong value region
    float4 *ring = smem4 + (size_t)sp.vrows_total * 8;                   // kGroups x 2 blocks x kBlk sorted entries
    float4 *gring = ring + kGroups * 2 * kBlk;                           // (kAhead+1) x CH x kWalkThreads private grad_out chunks
    int *cell_off = reinterpret_cast<int *>(gring + (kAhead + 1) * CH * kWalkThreads);  // kMaxCells + 8 counters (sorter only)
    int *hdr = cell_off + kMaxCells + 8;                                 // [2][4]: {E0, E} of the unit in buffer b
    int *warp_tot = hdr + 8;
    float2 *st_loc = reinterpret_cast<float2 *>(warp_tot + 32);          // sorter staging: locations of stage_q queries' samples
    float *st_att = reinterpret_cast<float *>(st_loc + sp.stage_q * P);  //                 and their attention weights
    const int tid = threadIdx.x;
    const int pix = sp.M * kD;
    const int qstride = sp.M * LP;  // samples between consecutive queries of one head
    const int n_mine = sp.base + ((int)blockIdx.x < sp.rem * sp.parts ? 1 : 0);
    float4 *my_scratch = scratch + (size_t)blockIdx.x * 2 * sp.scratch_stride;

    // CTA b runs whole units [b*base, (b+1)*base) and then, if it exists, piece b of the leftover units
    auto decode = [&](int ui) {
        Unit u;
        const bool piece = ui == sp.base;
        const long long g = piece ? (long long)gridDim.x * sp.base + blockIdx.x / sp.parts : (long long)blockIdx.x * sp.base + ui;
        u.part = piece ? (int)blockIdx.x % sp.parts : 0;
        u.parts = piece ? sp.parts : 1;
        const long long nm = g / sp.L;
        u.l = (int)(g - nm * sp.L);
        u.n = nm / sp.M;
        u.m = (int)(nm - u.n * sp.M);
        return u;
    };
    // value slab of unit ui: low end of the region for even ui, high end for odd ui
    auto vrow0 = [&](int ui, int l) { return (ui & 1) ? sp.vrows_total - sp.vrows[l] : 0; };
    // do the slabs of unit ui-1 and ui collide?  (then the sorter must wait for the walkers to leave ui-1 before its slab copy)
    auto collide = [&](int l_prev, int l_cur) { return sp.vrows[l_prev] + sp.vrows[l_cur] > sp.vrows_total; };

    if (tid < kSortThreads) {
        // ================================================= SORTER =====================================================
        if (SW == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int lane = tid & 31, wid = tid >> 5;
        const unsigned loc_s = (unsigned)__cvta_generic_to_shared(st_loc), att_s = (unsigned)__cvta_generic_to_shared(st_att);
        int seen = 0;  // units the walkers are known to have finished (their kBarEmpty arrivals consumed so far)
        int l_prev = 0;
        for (int ui = 0; ui < n_mine; ++ui) {
            const Unit u = decode(ui);
            const int b = ui & 1, l = u.l;
            const int H = sp.H[l], W = sp.W[l], cs = sp.cshift[l], cells = (H + 1) << cs, Wp = W + 2;
            const float Hf = (float)H, Wf = (float)W;
            const long long lvl_row = (u.n * sp.S + sp.start[l]) * (long long)pix + (long long)u.m * kD;
            const float *vbase = value + lvl_row;
            const long long meta0 = ((u.n * Lq) * sp.M + u.m) * (long long)LP + l * P;
            const float *locb = loc + 2 * meta0;
            const float *att = attn + meta0;
            float *gloc = grad_loc + 2 * meta0;
            float *gatt = grad_attn + meta0;
            // scratch list / header of buffer b are free once the walkers finished unit ui-2
            while (seen < ui - 1) {
                bar_sync(kBarEmpty0 + (seen & 1), kThreads);
                ++seen;
            }
            // stage the samples of queries [q0, q0+nq): locations + weights, one memory round trip for all of them
            auto stage = [&](int q0, int nq) {
                bar_sync(kBarSort, kSortThreads);  // every sorter thread is done with what the staging buffer held
                if (P_T == 4) {  // the four points of a (query, head, level) are 32 + 16 contiguous, 16-byte aligned bytes
                    for (int q = tid; q < nq; q += kSortThreads) {
                        const float *lsrc = locb + 2 * ((q0 + q) * qstride);
                        cp_async16(loc_s + q * 32, lsrc);
                        cp_async16(loc_s + q * 32 + 16, lsrc + 4);
                        cp_async16(att_s + q * 16, att + (q0 + q) * qstride);
                    }
                } else {
                    for (int i = tid; i < nq * P; i += kSortThreads) {
                        const int q = i / P, p = i - q * P;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(loc_s + i * 8), "l"(locb + 2 * ((q0 + q) * qstride + p)) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(att_s + i * 4), "l"(att + (q0 + q) * qstride + p) : "memory");
                    }
                }
                cp_commit();
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                bar_sync(kBarSort, kSortThreads);
            };
            const int nchunk = (Lq + sp.stage_q - 1) / sp.stage_q;
            bar_sync(kBarSort, kSortThreads);  // every sorter thread has left pass 2 of the previous unit (it still bumps cell_off)
            for (int i = tid; i <= cells; i += kSortThreads) cell_off[i] = 0;  // (stage() below starts with a sorter barrier)
            // ---- pass 1: histogram of cells; samples outside the map get their (exactly zero) gradients here -------------------
            for (int c = 0; c < nchunk; ++c) {
                const int q0 = c * sp.stage_q, nq = min(sp.stage_q, Lq - q0);
                stage(q0, nq);
                for (int i = tid; i < nq * P; i += kSortThreads) {
                    const float2 xy = st_loc[i];
                    const float h = __fmaf_rn(xy.y, Hf, -0.5f), w = __fmaf_rn(xy.x, Wf, -0.5f);
                    if (h > -1.f && w > -1.f && h < Hf && w < Wf) {
                        int y0, x0;
                        floor_magic(h, y0);
                        floor_magic(w, x0);
                        atomicAdd(&cell_off[((y0 + 1) << cs) + x0 + 1], 1);
                    } else if (u.part == 0) {  // cuh:352-356: nothing is accumulated for such a sample
                        const int q = P_T > 0 ? i / P_T : i / P, p = i - q * P;
                        const int s = (q0 + q) * qstride + p;
                        *reinterpret_cast<float2 *>(gloc + 2 * s) = make_float2(0.f, 0.f);
                        gatt[s] = 0.f;
                    }
                }
            }
            bar_sync(kBarSort, kSortThreads);
            // ---- exclusive scan of <= kMaxCells counters: kCPT consecutive cells per thread ------------------------------------
            {
                constexpr int kCPT = kMaxCells / kSortThreads;
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
                bar_sync(kBarSort, kSortThreads);
                int basev = incl - sum;
#pragma unroll
                for (int k = 0; k < kSortWarps; ++k)
                    if (k < wid) basev += warp_tot[k];
                int total = 0;
#pragma unroll
                for (int k = 0; k < kSortWarps; ++k) total += warp_tot[k];
#pragma unroll
                for (int k = 0; k < kCPT; ++k) {
                    const int i = kCPT * tid + k;
                    if (i < cells) cell_off[i] = basev;
                    basev += v[k];
                }
                if (tid == 0) cell_off[kMaxCells] = total;  // number of in-range samples
            }
            bar_sync(kBarSort, kSortThreads);
            if (tid == 0) {
                // a piece takes whole cells: its range starts at the first cell whose offset reaches Etot*part/parts
                const int Etot = cell_off[kMaxCells];
                auto cell_boundary = [&](int target) {
                    int lo = 0, hi = cells;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (cell_off[mid] >= target) hi = mid; else lo = mid + 1;
                    }
                    return lo < cells ? cell_off[lo] : Etot;
                };
                int E0 = 0, E = Etot;
                if (u.parts > 1) {
                    E0 = u.part == 0 ? 0 : cell_boundary((int)((long long)Etot * u.part / u.parts));
                    E = (u.part + 1 == u.parts ? Etot : cell_boundary((int)((long long)Etot * (u.part + 1) / u.parts))) - E0;
                }
                hdr[b * 4 + 0] = E0;
                hdr[b * 4 + 1] = E;
            }
            // (the boundaries are read from the start offsets, which pass 2 advances: both branches below start with a sorter barrier)
            // ---- pass 2: every in-range sample claims the next slot of its cell and writes its entry to the scratch list ----------
            float4 *ent = my_scratch + (size_t)b * sp.scratch_stride;
            for (int c = nchunk - 1; c >= 0; --c) {  // backwards: the last chunk is still staged from pass 1
                const int q0 = c * sp.stage_q, nq = min(sp.stage_q, Lq - q0);
                if (c != nchunk - 1) stage(q0, nq);
                else bar_sync(kBarSort, kSortThreads);
#pragma unroll 4
                for (int i = tid; i < nq * P; i += kSortThreads) {
                    const float2 xy = st_loc[i];
                    const float h = __fmaf_rn(xy.y, Hf, -0.5f), w = __fmaf_rn(xy.x, Wf, -0.5f);
                    if (h > -1.f && w > -1.f && h < Hf && w < Wf) {
                        int y0, x0;
                        const float lh = h - floor_magic(h, y0), lw = w - floor_magic(w, x0);
                        const int cell = ((y0 + 1) << cs) + x0 + 1;
                        const int pos = atomicAdd(&cell_off[cell], 1);
                        const int q = P_T > 0 ? i / P_T : i / P, p = i - q * P;
                        ent[pos] = make_float4(lh, lw, st_att[i], __int_as_float((cell << 20) | (p << 16) | (q0 + q)));
                    }
                }
            }
            // ---- value level slab -> shared memory, last: only this step may have to wait for the walkers to leave unit ui-1 ------
            while (seen < ui - 1 + ((ui >= 1 && collide(l_prev, l)) ? 1 : 0)) {
                bar_sync(kBarEmpty0 + (seen & 1), kThreads);
                ++seen;
            }
            l_prev = l;
            {
                float4 *Vu = V4 + (size_t)vrow0(ui, l) * 8;
                const unsigned Vu_s = (unsigned)__cvta_generic_to_shared(Vu);
                for (int i = tid; i < H * W * 8; i += kSortThreads) {
                    const int pxl = i >> 3, c = i & 7;
                    const int y = (int)__umulhi((unsigned)pxl, sp.wrecip[l]), x = pxl - y * W;
                    cp_async16(Vu_s + (unsigned)(((y + 1) * Wp + x + 1) * 8 + c) * 16, vbase + pxl * pix + c * 4);
                }
                cp_commit();
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int i = tid; i < Wp * 8; i += kSortThreads) {  // the one-pixel zero border
                    Vu[i] = z;
                    Vu[(H + 1) * Wp * 8 + i] = z;
                }
                for (int i = tid; i < H * 16; i += kSortThreads) {
                    const int y = (i >> 4) + 1, side = (i >> 3) & 1, c = i & 7;
                    Vu[(y * Wp + side * (W + 1)) * 8 + c] = z;
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __threadfence_block();
            bar_arrive(kBarFull0 + b, kThreads);  // unit ui is ready in buffer b
        }
        return;
    }

    // ===================================================== WALKERS =====================================================
    if (SW == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int wt = tid - kSortThreads;
    const int grp = wt / SPLIT;
    int j = wt % SPLIT;
    pin(j);
    int chf[CH];  // float offset of slot k inside a 32-channel row; the 8 lanes of an LDS.128 phase cover all 32 banks
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        chf[k] = (j * CH + ((k + grp) % CH)) * 4;
        pin(chf[k]);
    }
    unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring) + (unsigned)grp * (2 * kBlk * 16);
    unsigned gslot_s = (unsigned)__cvta_generic_to_shared(gring) + (unsigned)wt * 16;  // slot (stage, k): + (stage*CH + k) * kWalkThreads * 16
    pin(ring_s); pin(gslot_s);
    const unsigned V_region = (unsigned)__cvta_generic_to_shared(V4);
    const float2 z2 = make_float2(0.f, 0.f);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int ui = 0; ui < n_mine; ++ui) {
        const Unit u = decode(ui);
        const int b = ui & 1, l = u.l;
        const int H = sp.H[l], W = sp.W[l], cs = sp.cshift[l], Wp = W + 2;
        const float Hf = (float)H, Wf = (float)W;
        const long long lvl_row = (u.n * sp.S + sp.start[l]) * (long long)pix + (long long)u.m * kD;
        const long long meta0 = ((u.n * Lq) * sp.M + u.m) * (long long)LP + l * P;
        float *gloc = grad_loc + 2 * meta0;
        float *gatt = grad_attn + meta0;
        const float *gobase = grad_out + (u.n * Lq * sp.M + u.m) * (long long)kD;
        float *gbase = grad_value + lvl_row;
        const int o01 = pix, o10 = W * pix, o11 = (W + 1) * pix;
        unsigned V_s = V_region + (unsigned)vrow0(ui, l) * (kD * 4);
        pin(V_s);
        int qs = qstride, pixv = pix, Wv = W, Wpb = Wp * (kD * 4);
        pin(qs); pin(pixv); pin(Wv); pin(Wpb);
        // result stores: lane 0 grad_attn, lane 1 grad_loc.y, lane 2 grad_loc.x (SPLIT >= 4)
        float *res_base = j == 0 ? gatt : gloc + (j == 1 ? 1 : 0);
        int res_stride = j == 0 ? 1 : 2;
        float res_scale = j == 0 ? 1.f : (j == 1 ? Hf : Wf);
        pin(res_base); pin(res_stride); pin(res_scale); pin(gobase); pin(gbase);

        bar_sync(kBarFull0 + b, kThreads);  // the sorter finished this unit: entries in scratch, slab in shared memory
        const int E0 = hdr[b * 4 + 0], E = hdr[b * 4 + 1];
        const int e_lo = E0 + (int)((long long)E * grp / kGroups), e_hi = E0 + (int)((long long)E * (grp + 1) / kGroups);
        const int full = E / kGroups;
        int cnt = e_hi - e_lo;
        pin(cnt);
        const int steps = full + (full * kGroups != E ? 1 : 0);  // the same for every group of the CTA
        const float4 *ent_g = my_scratch + (size_t)b * sp.scratch_stride + e_lo;  // cp.async.cg: L2, written by this CTA's sorter
        pin(ent_g);

        // ring: entry i of the group's range sits at ring_s + (i & 15) * 16; a block is kBlk = 8 entries, lanes copy one (two) each
        auto copy_block = [&](int kb) {
            for (int t = j; t < kBlk; t += SPLIT) cp_async16(ring_s + (((kb * kBlk) + t) & 15) * 16, ent_g + kb * kBlk + t);
        };
        // grad_out chunks of the entry at `step` -> this lane's slot `stage`; nothing to copy past the group's last entry
        auto copy_g = [&](int step, int stage) {
            if (step < cnt) {
                const unsigned packed = lds1(ring_s + (step & 15) * 16 + 12);
                const float *gq = gobase + (packed & 0xffff) * pixv;
#pragma unroll
                for (int k = 0; k < CH; ++k) cp_async16(gslot_s + (unsigned)((stage * CH + k) * kWalkThreads) * 16, gq + chf[k]);
            }
        };
        __syncwarp();
        copy_block(0);
        copy_block(1);
        cp_commit();
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int d = 0; d < kAhead; ++d) {
            copy_g(d, d);
            cp_commit();
        }

        int cur = -1, inmask = 0;
        float *gp = gbase;
        f4p v00[CH], dW[CH], dH[CH], dd[CH];
        f4p a00[CH], a01[CH], a10[CH], a11[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) a00[k] = a01[k] = a10[k] = a11[k] = {z2, z2};
        auto flush = [&]() {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                float *gk = gp + chf[k];
                if (inmask & 1) red_add_v4(gk, a00[k].lo.x, a00[k].lo.y, a00[k].hi.x, a00[k].hi.y);
                if (inmask & 2) red_add_v4(gk + o01, a01[k].lo.x, a01[k].lo.y, a01[k].hi.x, a01[k].hi.y);
                if (inmask & 4) red_add_v4(gk + o10, a10[k].lo.x, a10[k].lo.y, a10[k].hi.x, a10[k].hi.y);
                if (inmask & 8) red_add_v4(gk + o11, a11[k].lo.x, a11[k].lo.y, a11[k].hi.x, a11[k].hi.y);
            }
        };
        // deferred per-sample tail: transposed butterfly over the group's lanes, then three plain stores
        float d_a = 0.f, d_h = 0.f, d_w = 0.f, d_att = 0.f;
        int d_idx = -1;
        auto finish = [&](float s_a, float s_h, float s_w, float att_v, int sidx) {
            float x = (j & 1) ? s_h : s_a;
            x += __shfl_xor_sync(0xffffffffu, (j & 1) ? s_a : s_h, 1);
            s_w += __shfl_xor_sync(0xffffffffu, s_w, 1);
            float keep = (j & 2) ? s_w : x;
            keep += __shfl_xor_sync(0xffffffffu, (j & 2) ? x : s_w, 2);
            if (SPLIT == 8) keep += __shfl_xor_sync(0xffffffffu, keep, 4);
            if (sidx >= 0 && j < 3) res_base[sidx * res_stride] = keep * res_scale * (j == 0 ? 1.f : att_v);
        };
        // one step: S = its index in the group's range, T = S & 7 (compile-time in the unrolled steady loop), ACTIVE = S < cnt
#define MSDA_PIPE_STEP(S, T, ACTIVE) \
        { \
            if ((T) == 0) { \
                __syncwarp();                       /* every lane is done with the ring block that is overwritten now */ \
                if ((S) > 0) copy_block(((S) >> 3) + 1); \
            } \
            copy_g((S) + kAhead, ((T) + kAhead) & 3); \
            cp_commit(); \
            asm volatile("cp.async.wait_group 3;" ::: "memory");  /* this step's grad_out chunk (and older copies) landed */ \
            if ((T) == 4) __syncwarp();             /* the ring block copied at T == 0 is complete in every lane */ \
            const float4 en_ = (ACTIVE) ? lds4(ring_s + ((S) & 15) * 16) : z4; \
            float4 gr_[CH]; \
_Pragma("unroll") \
            for (int k = 0; k < CH; ++k) gr_[k] = (ACTIVE) ? lds4(gslot_s + (unsigned)((((T) & 3) * CH + k) * kWalkThreads) * 16) : z4; \
            float s_a = 0.f, s_h = 0.f, s_w = 0.f, att_v = 0.f; \
            int sidx = -1; \
            if (ACTIVE) { \
                const int packed = __float_as_int(en_.w); \
                const int cell = packed >> 20; \
                sidx = (packed & 0xffff) * qs + ((packed >> 16) & 0xf); \
                att_v = en_.z; \
                if (cell != cur) { \
                    flush(); \
                    cur = cell; \
                    const int cy = cell >> cs, cx = cell & ((1 << cs) - 1); \
                    inmask = ((cy >= 1 && cx >= 1) ? 1 : 0) | ((cy >= 1 && cx <= Wv - 1) ? 2 : 0) | \
                             ((cy <= H - 1 && cx >= 1) ? 4 : 0) | ((cy <= H - 1 && cx <= Wv - 1) ? 8 : 0); \
                    gp = gbase + ((cy - 1) * Wv + cx - 1) * pixv; \
                    const unsigned vr = V_s + (unsigned)(cy * Wpb + cx * (kD * 4)); \
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
                const float lh = en_.x, lw = en_.y; \
                const float ah = (1.f - lh) * att_v, al = lh * att_v, hw = 1.f - lw; \
                const float2 lh2 = make_float2(lh, lh), lw2 = make_float2(lw, lw); \
                const float2 w1 = make_float2(ah * hw, ah * hw), w2 = make_float2(ah * lw, ah * lw); \
                const float2 w3 = make_float2(al * hw, al * hw), w4 = make_float2(al * lw, al * lw); \
                float2 pa = z2, ph = z2, pw = z2; \
_Pragma("unroll") \
                for (int k = 0; k < CH; ++k) { \
                    const f4p g = pk(gr_[k]); \
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
            if (DEFER) { \
                finish(d_a, d_h, d_w, d_att, d_idx); \
                d_a = s_a; d_h = s_h; d_w = s_w; d_att = att_v; d_idx = sidx; \
            } else { \
                finish(s_a, s_h, s_w, att_v, sidx); \
            } \
        }

        // steady blocks of kBlk steps: every group is active in every step (s < full <= cnt)
        int s = 0;
        for (; s + kBlk <= full; s += kBlk) {
            MSDA_PIPE_STEP(s + 0, 0, true)
            MSDA_PIPE_STEP(s + 1, 1, true)
            MSDA_PIPE_STEP(s + 2, 2, true)
            MSDA_PIPE_STEP(s + 3, 3, true)
            MSDA_PIPE_STEP(s + 4, 4, true)
            MSDA_PIPE_STEP(s + 5, 5, true)
            MSDA_PIPE_STEP(s + 6, 6, true)
            MSDA_PIPE_STEP(s + 7, 7, true)
        }
        // the remaining < kBlk + 1 steps (uniform count; a group may be inactive in the last one)
        for (; s < steps; ++s) {
            const int t = s & 7;
            const bool active = s < cnt;
            MSDA_PIPE_STEP(s, t, active)
        }
#undef MSDA_PIPE_STEP
        if (DEFER) {
            finish(d_a, d_h, d_w, d_att, d_idx);
            d_idx = -1;
        }
        if (cur >= 0) flush();
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // no stale copy may land during the next unit
        // tell the sorter this unit's buffers are free -- only when it will wait for that (see its two `seen` loops)
        bool tell = ui + 2 < n_mine;
        if (ui + 2 == n_mine) tell = collide(l, decode(ui + 1).l);
        if (tell) bar_arrive(kBarEmpty0 + b, kThreads);
    }
}

int pipe_sort_warps()
{
    static const int v = [] { const char *e = getenv("MSDA_PIPE_SORTW"); return e && atoi(e) == 2 ? 2 : 4; }();
    return v;
}
int pipe_split()
{
    static const int v = [] { const char *e = getenv("MSDA_PIPE_SPLIT"); return e && atoi(e) == 4 ? 4 : 8; }();
    return v;
}
bool pipe_defer()
{
    static const bool v = [] { const char *e = getenv("MSDA_PIPE_DEFER"); return !(e && atoi(e) == 0); }();
    return v;
}

bool make_params(const Problem &p, PipeParams &sp, size_t &smem)
{
    if (!(p.dtype == 0 && p.D == kD && p.have_host && p.L <= kMaxLevels)) return false;
    if (p.Lq < 1 || p.Lq > 65535 || p.P > 16 || (long long)p.Lq * p.P > kMaxSamples) return false;
    sp.L = p.L; sp.P = p.P; sp.S = p.S; sp.M = p.M; sp.Lq = p.Lq;
    int vmax = 0;
    for (int l = 0; l < p.L; ++l) {
        sp.H[l] = p.table.H[l];
        sp.W[l] = p.table.W[l];
        sp.start[l] = p.table.start[l];
        int cs = 0;
        while ((1 << cs) < sp.W[l] + 1) ++cs;
        sp.cshift[l] = cs;
        if (((long long)(sp.H[l] + 1) << cs) > kMaxCells) return false;
        if ((long long)(sp.H[l] + 2) * (sp.W[l] + 2) * p.M * kD >= (1LL << 31)) return false;
        if ((long long)sp.H[l] * sp.W[l] >= 65536) return false;
        sp.wrecip[l] = (unsigned)((0x100000000ULL + sp.W[l] - 1) / sp.W[l]);
        sp.vrows[l] = (sp.H[l] + 2) * (sp.W[l] + 2);
        if (sp.vrows[l] > vmax) vmax = sp.vrows[l];
    }
    const int walk_threads = kThreads - pipe_sort_warps() * 32, groups = walk_threads / pipe_split();
    const int ch = 8 / pipe_split();
    const size_t fixed = (size_t)groups * 2 * kBlk * 16 + (size_t)(kAhead + 1) * ch * walk_threads * 16 + (kMaxCells + 8 + 8 + 32) * sizeof(int);
    const size_t budget = (size_t)p.smem_optin;
    const size_t per_q = (size_t)p.P * 12;  // staged bytes per query: P locations + P weights
    if ((size_t)vmax * 128 + fixed + 64 * per_q > budget) return false;
    // Ping-pong value region: the largest sum of two consecutive levels' slabs that still leaves room to stage half of a
    // unit's samples (a slab copy then waits for the walkers only on the transitions that do not fit); at least the largest slab.
    int half_q = (p.Lq + 1) / 2;
    if (half_q > p.Lq) half_q = p.Lq;
    int v_total = vmax;
    for (int l = 0; l < p.L; ++l) {
        const int s = sp.vrows[l] + sp.vrows[(l + 1) % p.L];
        if (s > v_total && (size_t)s * 128 + fixed + (size_t)half_q * per_q <= budget) v_total = s;
    }
    sp.vrows_total = v_total;
    const size_t left = budget - fixed - (size_t)v_total * 128;
    int stage_q = (int)(left / per_q);
    stage_q &= ~3;  // keeps the weights' staging area 16-byte aligned
    if (stage_q >= p.Lq) stage_q = (p.Lq + 3) & ~3;
    if (stage_q < 4) return false;
    sp.stage_q = stage_q;
    sp.scratch_stride = p.Lq * p.P + kScratchPad;
    sp.units = (long long)p.N * p.M * p.L;
    smem = (size_t)sp.vrows_total * 128 + fixed + (size_t)stage_q * per_q;
    return smem <= budget;
}

}  // namespace

bool pipe32_supported(const Problem &p)
{
    PipeParams sp;
    size_t smem;
    return make_params(p, sp, smem);
}

size_t bwd_pipe32_workspace(int dtype, int D, int Lq, int P)
{
    if (dtype != 0 || D != kD || Lq < 1 || (long long)Lq * P > kMaxSamples) return 0;
    return (size_t)kMaxCtas * 2 * ((size_t)Lq * P + kScratchPad) * sizeof(float4);
}

cudaError_t bwd_pipe32(const Problem &p, const BackwardIO &io, cudaStream_t st)
{
    PipeParams sp;
    size_t smem;
    if (!make_params(p, sp, smem)) return cudaErrorInvalidValue;
    const long long n_val = (long long)p.N * p.S * p.M * p.D;
    if (n_val > 0) {
        cudaError_t e = cudaMemsetAsync(io.grad_value, 0, sizeof(float) * (size_t)n_val, st);
        if (e != cudaSuccess) return e;
    }
    if (sp.units == 0 || p.Lq == 0) return cudaSuccess;
    if (!io.workspace || io.workspace_bytes < bwd_pipe32_workspace(p.dtype, p.D, p.Lq, p.P)) return cudaErrorInvalidValue;
    const bool s4 = pipe_split() == 4, df = pipe_defer(), w4 = pipe_sort_warps() == 4;
    auto pick = [&](auto p_t) {
        constexpr int PT = decltype(p_t)::value;
        return w4 ? (s4 ? (df ? bwd_pipe32_kernel<PT, 4, true, 4> : bwd_pipe32_kernel<PT, 4, false, 4>)
                        : (df ? bwd_pipe32_kernel<PT, 8, true, 4> : bwd_pipe32_kernel<PT, 8, false, 4>))
                  : (s4 ? (df ? bwd_pipe32_kernel<PT, 4, true, 2> : bwd_pipe32_kernel<PT, 4, false, 2>)
                        : (df ? bwd_pipe32_kernel<PT, 8, true, 2> : bwd_pipe32_kernel<PT, 8, false, 2>));
    };
    auto kern = p.P == 4 ? pick(std::integral_constant<int, 4>()) : pick(std::integral_constant<int, 0>());
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = p.sm_count < kMaxCtas ? p.sm_count : kMaxCtas;
    if (sp.units < grid) grid = (int)sp.units;
    sp.base = (int)(sp.units / grid);
    sp.rem = (int)(sp.units - (long long)sp.base * grid);
    sp.parts = sp.rem > 0 ? grid / sp.rem : 1;
    if (sp.parts > 4) sp.parts = 4;
    if (sp.parts < 1) sp.parts = 1;
    kern<<<grid, kThreads, smem, st>>>((const float *)p.value, (const float *)p.loc, (const float *)p.attn, (const float *)io.grad_out,
                                       (float *)io.grad_value, (float *)io.grad_loc, (float *)io.grad_attn, (float4 *)io.workspace, sp);
    count_launch();
    return cudaGetLastError();
}

}  // namespace msda
