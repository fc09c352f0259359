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
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "msda_kernels.h"
#include "msda_tma.cuh"

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
    float rw[kMaxLevels], rh[kMaxLevels];  // 1/W, 1/H (fused prologue)
    unsigned wrecip[kMaxLevels];  // ceil(2^32 / W): pixel / W == umulhi(pixel, wrecip) for pixel < 2^16
    unsigned precip;          // ceil(2^32 / (3*L*P/4)) for the metadata staging index split
    int padded_rows;          // total rows of the padded slab
    int wstage;               // every warp stages its own rows' metadata: the per-pass CTA barrier becomes a warp barrier
    int fused;                // loc = raw offsets, attn = raw logits: transform the staged rows in place (needs STAGE)
    int qpp;                  // queries per pass
    int pps;                  // passes per slab
    long long units;          // N*M*pps
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
using bf16 = __nv_bfloat16;

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}

// 8 consecutive bf16 (16 bytes) -> two float4
__device__ __forceinline__ void bf16x8_to_f32(const uint4 &r, float4 &a, float4 &b)
{
    const float2 p0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
    const float2 p1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
    const float2 p2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.z));
    const float2 p3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.w));
    a = make_float4(p0.x, p0.y, p1.x, p1.y);
    b = make_float4(p2.x, p2.y, p3.x, p3.y);
}

// copy value[n, :, m, :] into the padded fp32 slab (interior only; the border was zeroed once)
__device__ __forceinline__ void load_slab(float4 *slab, const float *__restrict__ value, const SlabParams &sp, int nm)
{
    const long long n = nm / sp.M;  // (32-bit division; 64-bit only for the element offsets below)
    const int m = nm - (int)n * sp.M;
    const float *src = value + ((n * sp.S) * sp.M + m) * kD;
    const int pixf = sp.M * kD;
    for (int l = 0; l < sp.L; ++l) {
        const int W = sp.W[l], cnt = sp.H[l] * W * 8;
        const float *lsrc = src + (long long)sp.start[l] * pixf;
        float4 *ldst = slab + (size_t)(sp.prow0[l] + (W + 2) + 1) * 8;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int pix = i >> 3, c = i & 7;
            const int y = (int)__umulhi((unsigned)pix, sp.wrecip[l]), x = pix - y * W;
            cp_async16(ldst + (size_t)(y * (W + 2) + x) * 8 + c, lsrc + (long long)pix * pixf + c * 4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per level: the gather of level l waits only for it
    }
}

// bf16 values: rows are 64 bytes in HBM; they are widened to fp32 on the way into the slab (the gather itself is unchanged)
__device__ __forceinline__ void load_slab(float4 *slab, const bf16 *__restrict__ value, const SlabParams &sp, int nm)
{
    const long long n = nm / sp.M;
    const int m = nm - (int)n * sp.M;
    const bf16 *src = value + ((n * sp.S) * sp.M + m) * kD;
    const int pixf = sp.M * kD;
    for (int l = 0; l < sp.L; ++l) {
        const int W = sp.W[l], cnt = sp.H[l] * W * 4;
        const bf16 *lsrc = src + (long long)sp.start[l] * pixf;
        float4 *ldst = slab + (size_t)(sp.prow0[l] + (W + 2) + 1) * 8;
        // four independent 16-byte loads in flight per thread before the first conversion / store (the plain loop serialises
        // one L2 round trip per iteration: the shared-memory stores may alias the loads as far as the compiler knows)
        for (int i0 = threadIdx.x; i0 < cnt; i0 += 4 * blockDim.x) {
            uint4 raw[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = i0 + r * blockDim.x;
                if (i < cnt) raw[r] = __ldg(reinterpret_cast<const uint4 *>(lsrc + (long long)(i >> 2) * pixf + (i & 3) * 8));
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = i0 + r * blockDim.x;
                if (i < cnt) {
                    const int pix = i >> 2, c = i & 3;
                    const int y = (int)__umulhi((unsigned)pix, sp.wrecip[l]), x = pix - y * W;
                    float4 a, b;
                    bf16x8_to_f32(raw[r], a, b);
                    float4 *d = ldst + (size_t)(y * (W + 2) + x) * 8 + c * 2;
                    d[0] = a;
                    d[1] = b;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // (empty group: keeps the group arithmetic identical to the fp32 path)
    }
}

__device__ __forceinline__ float2 to_f2(float2 v) { return v; }
__device__ __forceinline__ float2 to_f2(__nv_bfloat162 v) { return __bfloat1622float2(v); }
__device__ __forceinline__ float to_f1(float v) { return v; }
__device__ __forceinline__ float to_f1(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void store4(bf16 *p, float4 v)
{
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<unsigned *>(&lo);
    r.y = *reinterpret_cast<unsigned *>(&hi);
    *reinterpret_cast<uint2 *>(p) = r;
}
template <typename T> struct Pair;
template <> struct Pair<float> { using type = float2; };
template <> struct Pair<bf16> { using type = __nv_bfloat162; };

// wait until at most n cp.async groups of this thread are pending (n is small; the instruction needs an immediate)
__device__ __forceinline__ void cp_async_wait_pending(int n)
{
    switch (n) {
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    }
}

// floor for |x| < 2^22 without the XU pipe: adding 1.5*2^23 with round-down leaves floor(x) in the low mantissa bits
__device__ __forceinline__ void floor_split(float x, float &fl, int &i)
{
    const float t = __fadd_rd(x, 12582912.f);
    fl = t - 12582912.f;
    i = __float_as_int(t) - 0x4B400000;
}

// Stage the sampling locations + attention weights of one pass (queries [q0, q0+cnt) of slab nm) into shared memory:
// per query 2*LP location floats followed by LP weights (row of 3*LP floats).  In global memory one (query, head)
// owns 8*LP contiguous bytes of locations and 4*LP of weights at a stride of M times that, so 16-byte cp.async
// pieces issued by consecutive lanes touch each 128-byte line once, instead of once per sample per query.
template <typename TM>
__device__ __forceinline__ void stage_meta(TM *meta, const TM *__restrict__ loc, const TM *__restrict__ attn,
                                           const SlabParams &sp, int LP, int nm, int q0, int cnt,
                                           TM *refs, const TM *__restrict__ ref, int rows_per_warp)
{
    const long long n = nm / sp.M;
    const int m = nm - (int)n * sp.M;
    const int pieces = (3 * LP) >> 2, loc_pieces = LP >> 1;  // pieces of 4 elements (16 bytes of fp32, 8 bytes of bf16)
    // rows_per_warp > 0: every warp stages only the rows its own lanes will read (no CTA barrier needed per pass);
    // rows_per_warp == 0: the whole CTA stages the whole pass
    int r0 = 0, nrows = cnt, t = threadIdx.x, nt = blockDim.x;
    if (rows_per_warp > 0) {
        r0 = (threadIdx.x >> 5) * rows_per_warp;
        nrows = min(max(cnt - r0, 0), rows_per_warp);
        t = threadIdx.x & 31;
        nt = 32;
    }
    const int total = nrows * pieces;
    const long long qm0 = (n * sp.Lq + q0 + r0) * sp.M + m;
    const TM *lbase = loc + qm0 * (2 * LP), *abase = attn + qm0 * LP;
    const int lstride = sp.M * 2 * LP, astride = sp.M * LP;
    TM *mrow = meta + r0 * (3 * LP);
    for (int i = t; i < total; i += nt) {
        const int ql = (int)__umulhi((unsigned)i, sp.precip), pc = i - ql * pieces;
        const TM *src = pc < loc_pieces ? lbase + ql * lstride + pc * 4 : abase + ql * astride + (pc - loc_pieces) * 4;
        if (sizeof(TM) == 4) cp_async16(mrow + ql * (3 * LP) + pc * 4, src);
        else cp_async8(mrow + ql * (3 * LP) + pc * 4, src);
    }
    if (refs != nullptr) {  // fused prologue: the rows' reference points, [rows][L][2] contiguous in global memory
        const TM *rsrc = ref + ((n * sp.Lq + q0 + r0) * sp.L) * 2;
        TM *rdst = refs + r0 * sp.L * 2;
        for (int i = t; i < nrows * sp.L; i += nt) {
            const unsigned d = (unsigned)__cvta_generic_to_shared(rdst + i * 2);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(rsrc + i * 2) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// packed fp32x2 (FFMA2 on sm_100a): halves the issue slots of the interpolation math
struct f4p {
    float2 lo, hi;
};
__device__ __forceinline__ float4 lds4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void pin(int &x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(unsigned &x) { asm volatile("" : "+r"(x)); }

template <typename TV, typename TM, int P_T, bool STAGE, int SPLIT>
__global__ void __launch_bounds__(SPLIT == 4 ? 768 : 512, 1)
fwd_slab32_kernel(const TV *__restrict__ value, const TM *__restrict__ loc, const TM *__restrict__ attn,
                  TV *__restrict__ out, const TM *__restrict__ ref, TM *__restrict__ attn_out, const unsigned char *__restrict__ mask,
                  const __grid_constant__ SlabParams sp)
{
    constexpr int CH = 8 / SPLIT;  // 16-byte chunks (4 channels) per lane
    extern __shared__ float4 slab[];
    const int P = P_T > 0 ? P_T : sp.P;
    const int LP = sp.L * P;
    TM *meta0 = reinterpret_cast<TM *>(slab + (size_t)sp.padded_rows * 8);  // two staging buffers of qpp rows
    const int meta_stride = sp.qpp * 3 * LP;
    // fused prologue: two buffers of qpp x L reference points after the metadata buffers
    TM *refs0 = meta0 + 2 * meta_stride;
    const int ref_stride = sp.qpp * sp.L * 2;
    const bool fused = STAGE && sizeof(TM) == 4 && sp.fused != 0;
    const int rpw = sp.wstage ? 32 / SPLIT : 0;  // (query, head) rows per warp
    for (int l = 0; l < sp.L; ++l) {  // zero the one-pixel border of every level once; slab loads only ever write interiors
        const int H = sp.H[l], Wp = sp.W[l] + 2;
        float4 *lv = slab + (size_t)sp.prow0[l] * 8;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = threadIdx.x; i < Wp * 8; i += blockDim.x) {
            lv[i] = z;
            lv[(H + 1) * Wp * 8 + i] = z;
        }
        for (int i = threadIdx.x; i < H * 16; i += blockDim.x) {
            const int y = (i >> 4) + 1, side = (i >> 3) & 1, c = i & 7;
            lv[(y * Wp + side * (Wp - 1)) * 8 + c] = z;
        }
    }
    // (unit and slab indices fit 32 bits -- checked on the host; every thread decodes them every pass, and 64-bit divisions cost
    // ~100 instructions each)
    const int u_lo = (int)(sp.units * blockIdx.x / gridDim.x), u_hi = (int)(sp.units * (blockIdx.x + 1) / gridDim.x);
    int ql = threadIdx.x / SPLIT;
    const int part = threadIdx.x % SPLIT;
    pin(ql);
    // lane (ql, part) owns chunks [part*CH, part*CH+CH); its k-th load reads chunk part*CH + ((k + ql) % CH), so the 8 lanes of
    // every LDS.128 phase touch 8 different chunks = all 32 banks, whatever rows they gather: conflict-free
    unsigned off[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        off[k] = (part * CH + ((k + ql) % CH)) * 16;
        pin(off[k]);
    }
    unsigned slab_s = (unsigned)__cvta_generic_to_shared(slab);
    pin(slab_s);
    int resident = -1;
    auto pass_range = [&](int u, int &nm, int &q0, int &cnt) {
        nm = u / sp.pps;
        q0 = (u - nm * sp.pps) * sp.qpp;
        cnt = min(sp.qpp, sp.Lq - q0);
        if (cnt < 0) cnt = 0;
    };
    if (STAGE && u_lo < u_hi) {
        int nm, q0, cnt;
        pass_range(u_lo, nm, q0, cnt);
        stage_meta(meta0, loc, attn, sp, LP, nm, q0, cnt, fused ? refs0 : nullptr, ref, rpw);
    }
    for (int u = u_lo; u < u_hi; ++u) {
        int nm, q0, cnt;
        pass_range(u, nm, q0, cnt);
        const TM *meta = meta0 + ((u - u_lo) & 1) * meta_stride;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const bool fresh = nm != resident;  // a new slab: its levels arrive as separate cp.async groups
        // staged metadata visible.  A new slab (and the first pass, after the zero fill) needs the whole CTA to have left the
        // previous pass; with warp-private staging a pass on the resident slab only needs the warp itself.
        if (fresh || rpw == 0) __syncthreads();
        else __syncwarp();
        if (fresh) {
            load_slab(slab, value, sp, nm);
            resident = nm;
        }
        if (STAGE) {  // prefetch the next pass's metadata while this one computes (always one group, possibly empty)
            if (u + 1 < u_hi) {
                int nm2, q02, cnt2;
                pass_range(u + 1, nm2, q02, cnt2);
                stage_meta(meta0 + ((u + 1 - u_lo) & 1) * meta_stride, loc, attn, sp, LP, nm2, q02, cnt2,
                           fused ? refs0 + ((u + 1 - u_lo) & 1) * ref_stride : nullptr, ref, rpw);
            } else {
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
        const bool live = ql < cnt;
        const int q = q0 + ql;
        const long long n = nm / sp.M;
        const int m = nm - (int)n * sp.M;
        const long long qm = (n * sp.Lq + q) * sp.M + m;
        using TM2 = typename Pair<TM>::type;
        const TM2 *loc2 = STAGE ? reinterpret_cast<const TM2 *>(meta + ql * 3 * LP) : reinterpret_cast<const TM2 *>(loc) + qm * LP;
        const TM *att = STAGE ? meta + ql * 3 * LP + 2 * LP : attn + qm * LP;
        if (fused) {
            // Module prologue in place on the staged row of (query, head) (ops/modules/ms_deform_attn.py:102-116).  The two lanes of
            // the pair run ONE instruction stream on the two halves of the row (no divergence): half of the L*P logits each for the
            // softmax (max and sum exchanged with one shuffle each; ex2-based exp, one reciprocal), half of the samples each for
            // location = reference point + offset * (1/W_l, 1/H_l) (one FMA).  Against the reference's expf / division kernels that
            // is a difference of a few ulp in the weights and <= 1 ulp in the locations.
            float *row = reinterpret_cast<float *>(const_cast<TM *>(meta)) + ql * 3 * LP;
            const int hl = LP >> 1;                      // LP % 4 == 0 (STAGE)
            float *lg = row + 2 * LP + part * hl;
            const float *rf = reinterpret_cast<const float *>(refs0 + ((u - u_lo) & 1) * ref_stride) + ql * sp.L * 2;
            if (LP == 12 && P == 4) {
                // the production shape (3 levels x 4 points): the lane's half row lives in registers, vector loads / stores only
                float2 g0 = make_float2(0.f, 0.f), g1 = g0, g2 = g0;
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0, o2 = o0;
                float4 *xy4 = reinterpret_cast<float4 *>(row) + part * 3;
                if (live) {
                    g0 = reinterpret_cast<const float2 *>(lg)[0]; g1 = reinterpret_cast<const float2 *>(lg)[1]; g2 = reinterpret_cast<const float2 *>(lg)[2];
                    o0 = xy4[0]; o1 = xy4[1]; o2 = xy4[2];
                }
                float mx = fmaxf(fmaxf(fmaxf(g0.x, g0.y), fmaxf(g1.x, g1.y)), fmaxf(g2.x, g2.y));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                g0.x = __expf(g0.x - mx); g0.y = __expf(g0.y - mx); g1.x = __expf(g1.x - mx);
                g1.y = __expf(g1.y - mx); g2.x = __expf(g2.x - mx); g2.y = __expf(g2.y - mx);
                float sum = (g0.x + g0.y) + (g1.x + g1.y) + (g2.x + g2.y);
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                if (live) {
                    const float inv = __frcp_rn(sum);
                    g0.x *= inv; g0.y *= inv; g1.x *= inv; g1.y *= inv; g2.x *= inv; g2.y *= inv;
                    reinterpret_cast<float2 *>(lg)[0] = g0; reinterpret_cast<float2 *>(lg)[1] = g1; reinterpret_cast<float2 *>(lg)[2] = g2;
                    float2 *ao = reinterpret_cast<float2 *>(reinterpret_cast<float *>(attn_out) + qm * LP + part * hl);
                    ao[0] = g0; ao[1] = g1; ao[2] = g2;
                    // samples part*6 .. part*6+5: part 0 -> levels 0,0,0,0,1,1; part 1 -> levels 1,1,2,2,2,2
                    const int la = part, lb = part + 1;
                    const float wa = sp.rw[la], ha = sp.rh[la], wb = sp.rw[lb], hb = sp.rh[lb];
                    const float xa = rf[2 * la], ya = rf[2 * la + 1], xb = rf[2 * lb], yb = rf[2 * lb + 1];
                    const bool p0 = part == 0;  // the middle float4 (samples 2, 3 of the half) belongs to level a for part 0, b for part 1
                    o0.x = __fmaf_rn(o0.x, wa, xa); o0.y = __fmaf_rn(o0.y, ha, ya); o0.z = __fmaf_rn(o0.z, wa, xa); o0.w = __fmaf_rn(o0.w, ha, ya);
                    o1.x = __fmaf_rn(o1.x, p0 ? wa : wb, p0 ? xa : xb); o1.y = __fmaf_rn(o1.y, p0 ? ha : hb, p0 ? ya : yb);
                    o1.z = __fmaf_rn(o1.z, p0 ? wa : wb, p0 ? xa : xb); o1.w = __fmaf_rn(o1.w, p0 ? ha : hb, p0 ? ya : yb);
                    o2.x = __fmaf_rn(o2.x, wb, xb); o2.y = __fmaf_rn(o2.y, hb, yb); o2.z = __fmaf_rn(o2.z, wb, xb); o2.w = __fmaf_rn(o2.w, hb, yb);
                    xy4[0] = o0; xy4[1] = o1; xy4[2] = o2;
                }
            } else {
                float mx = -3.0e38f;
                if (live)
                    for (int i = 0; i < hl; ++i) mx = fmaxf(mx, lg[i]);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                float sum = 0.f;
                if (live)
                    for (int i = 0; i < hl; ++i) {
                        const float e = __expf(lg[i] - mx);
                        lg[i] = e;
                        sum += e;
                    }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                if (live) {
                    const float inv = __frcp_rn(sum);
                    float2 *ao = reinterpret_cast<float2 *>(reinterpret_cast<float *>(attn_out) + qm * LP + part * hl);
                    for (int i = 0; i < hl; i += 2) {
                        const float2 a2 = make_float2(lg[i] * inv, lg[i + 1] * inv);
                        *reinterpret_cast<float2 *>(lg + i) = a2;
                        ao[i >> 1] = a2;
                    }
                    float2 *xy = reinterpret_cast<float2 *>(row) + part * hl;
                    for (int i = 0; i < hl; ++i) {
                        const int l = (part * hl + i) / P;
                        float2 v = xy[i];
                        v.x = __fmaf_rn(v.x, sp.rw[l], rf[2 * l]);
                        v.y = __fmaf_rn(v.y, sp.rh[l], rf[2 * l + 1]);
                        xy[i] = v;
                    }
                }
            }
            if (rpw == 0) __syncthreads();  // both lanes of a pair read the whole row below
            else __syncwarp();
        }
        f4p acc[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) acc[k].lo = acc[k].hi = make_float2(0.f, 0.f);
        for (int l = 0; l < sp.L; ++l) {
            if (fresh) {  // level l of the new slab must have landed (levels after it and the metadata group may still be in flight)
                cp_async_wait_pending(sp.L - 1 - l + (STAGE ? 1 : 0));
                __syncthreads();
                if (fused && mask != nullptr) {  // padding mask folded in (ms_deform_attn.py:103-104): masked pixels read as zero rows
                    const int W = sp.W[l], npix = sp.H[l] * W;
                    const unsigned char *mk = mask + (long long)(nm / sp.M) * sp.S + sp.start[l];
                    float4 *lv = slab + (size_t)(sp.prow0[l] + (W + 2) + 1) * 8;
                    for (int pix = threadIdx.x; pix < npix; pix += blockDim.x) {
                        if (mk[pix]) {
                            const int y = (int)__umulhi((unsigned)pix, sp.wrecip[l]), x = pix - y * W;
                            float4 *row = lv + (size_t)(y * (W + 2) + x) * 8;
#pragma unroll
                            for (int c = 0; c < 8; ++c) row[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    __syncthreads();
                }
            }
            if (!live) continue;
            const float Hf = (float)sp.H[l], Wf = (float)sp.W[l];
            const float hmax = sp.hmax[l], wmax = sp.wmax[l];
            const int pitch = sp.W[l] + 2;
            const unsigned lvl_s = slab_s + (unsigned)(sp.prow0[l] + pitch + 1) * kRowBytes;  // address of pixel (0,0)
            const unsigned pitch_b = (unsigned)pitch * kRowBytes;
#pragma unroll
            for (int pp = 0; pp < (P_T > 0 ? P_T : 1); ++pp) {
                for (int p = (P_T > 0 ? pp : 0); p < (P_T > 0 ? pp + 1 : P); ++p) {
                    const int s = l * P + p;
                    const float2 xy = to_f2(STAGE ? loc2[s] : __ldg(loc2 + s));
                    float a = to_f1(STAGE ? att[s] : __ldg(att + s));
                    const float h = __fmaf_rn(xy.y, Hf, -0.5f), w = __fmaf_rn(xy.x, Wf, -0.5f);
                    const bool ok = h > -1.f && w > -1.f && h < Hf && w < Wf;
                    a = ok ? a : 0.f;
                    const float hc = fminf(fmaxf(h, -1.f), hmax), wc = fminf(fmaxf(w, -1.f), wmax);
                    float fh, fw;
                    int y0, x0;
                    floor_split(hc, fh, y0);
                    floor_split(wc, fw, x0);
                    const float lh = hc - fh, lw = wc - fw;
                    const float ahh = a * (1.f - lh), alh = a * lh, hw = 1.f - lw;
                    const float w1s = ahh * hw, w2s = ahh * lw, w3s = alh * hw, w4s = alh * lw;
                    const float2 w1 = make_float2(w1s, w1s), w2 = make_float2(w2s, w2s), w3 = make_float2(w3s, w3s), w4 = make_float2(w4s, w4s);
                    const unsigned r00 = lvl_s + (unsigned)(y0 * pitch + x0) * kRowBytes;
                    const unsigned r10 = r00 + pitch_b;
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        const float4 v1 = lds4(r00 + off[k]), v2 = lds4(r00 + off[k] + kRowBytes);
                        const float4 v3 = lds4(r10 + off[k]), v4 = lds4(r10 + off[k] + kRowBytes);
                        acc[k].lo = __ffma2_rn(w1, make_float2(v1.x, v1.y), __ffma2_rn(w2, make_float2(v2.x, v2.y),
                                    __ffma2_rn(w3, make_float2(v3.x, v3.y), __ffma2_rn(w4, make_float2(v4.x, v4.y), acc[k].lo))));
                        acc[k].hi = __ffma2_rn(w1, make_float2(v1.z, v1.w), __ffma2_rn(w2, make_float2(v2.z, v2.w),
                                    __ffma2_rn(w3, make_float2(v3.z, v3.w), __ffma2_rn(w4, make_float2(v4.z, v4.w), acc[k].hi))));
                    }
                }
            }
        }
        if (!live) continue;
        TV *o = out + qm * kD;
#pragma unroll
        for (int k = 0; k < CH; ++k) store4(o + (off[k] >> 2), make_float4(acc[k].lo.x, acc[k].lo.y, acc[k].hi.x, acc[k].hi.y));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------------------
// TMA-pipelined slab kernel (float32 values and metadata): the default forward.
//
// Same mapping, metadata staging and scheduling as fwd_slab32_kernel above.  What changes is how slabs move:
//   * the last warp is a PRODUCER: one elected lane issues one cp.async.bulk.tensor per level (a (H+2) x (W+2) pixel box at
//     (-1, -1); the out-of-bounds border arrives as zeros) and goes back to waiting; nobody computes addresses, nobody zero-fills;
//   * full[l] / empty[l] mbarriers per level hand the level regions back and forth.  Consumers release level l of the resident
//     slab as soon as their last pass is past it, so the next slab's small levels land while the 28x28 level is still being
//     gathered, and its 28x28 level lands under the first pass's small levels -- the slab switch that the cp.async version
//     serialises (a CTA barrier, 158 KB at the per-SM fill rate, another barrier per level) mostly disappears: 82.3 -> 73.9 us.
// (Round 2 also built the backward's grad_sampling_loc / grad_attn_weight as a second mode of this kernel -- a by-query gather next
// to a grad_value-only sort-and-segment kernel.  Parity-green, but 110 + 132 us against 206 us for msda_sorted32.cu, which does
// both in one walk with fewer instructions in total: removed again, measurements in profiles/r04_split_backward_experiments.txt.)
constexpr int kFwdTmaThreads = 512;  // including the producer warp

template <int P_T>
__global__ void __launch_bounds__(kFwdTmaThreads, 1)
slab32_tma_kernel(const float *__restrict__ loc, const float *__restrict__ attn, float *__restrict__ out,
                  const float *__restrict__ ref, float *__restrict__ attn_out, const unsigned char *__restrict__ mask,
                  const __grid_constant__ SlabParams sp, const __grid_constant__ TmaMaps maps)
{
    constexpr int CH = 4;  // SPLIT = 2: a lane owns four 16-byte chunks (16 channels)
    extern __shared__ __align__(128) float4 slab[];
    const int P = P_T > 0 ? P_T : sp.P;
    const int LP = sp.L * P;
    float *meta0 = reinterpret_cast<float *>(slab + (size_t)sp.padded_rows * 8);  // two staging buffers of qpp rows
    const int meta_stride = sp.qpp * 3 * LP;
    float *refs0 = meta0 + 2 * meta_stride;
    const int ref_stride = sp.qpp * sp.L * 2;
    const bool fused = sp.fused != 0;
    const int ref_floats = fused ? 2 * ref_stride : 0;
    const unsigned bars = (unsigned)__cvta_generic_to_shared(refs0 + ref_floats);  // full[l] at bars + 8*l, empty[l] at bars + 8*(4 + l)
    const int nwarps = blockDim.x >> 5, ncw = nwarps - 1, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u_lo = (int)(sp.units * blockIdx.x / gridDim.x), u_hi = (int)(sp.units * (blockIdx.x + 1) / gridDim.x);
    if (u_lo >= u_hi) return;
    const int nm_first = u_lo / sp.pps, nm_last = (u_hi - 1) / sp.pps;
    if (threadIdx.x == 0) {
        for (int l = 0; l < sp.L; ++l) {
            mbar_init(bars + 8 * l, 1);
            mbar_init(bars + 8 * (kTmaMaxLevels + l), ncw);
        }
        mbar_fence_init();
    }
    __syncthreads();
    unsigned slab_s = (unsigned)__cvta_generic_to_shared(slab);

    if (wid == ncw) {  // ---- producer warp ------------------------------------------------------------------------------------
        if (lane == 0) {
            for (int l = 0; l < sp.L; ++l) tma_prefetch_map(&maps.level[l]);
            int it = 0;
            for (int nm = nm_first; nm <= nm_last; ++nm, ++it) {
                const int n = nm / sp.M, m = nm - n * sp.M;
                for (int l = 0; l < sp.L; ++l) {
                    if (it > 0) mbar_wait(bars + 8 * (kTmaMaxLevels + l), (unsigned)((it - 1) & 1));  // consumers are past level l of the previous slab
                    const unsigned bytes = (unsigned)((sp.H[l] + 2) * (sp.W[l] + 2)) * kRowBytes;
                    mbar_arrive_expect_tx(bars + 8 * l, bytes);
                    // (L2 evict-first: a slab is read by one or two CTAs and never again -- DRAM reads 116 -> 105 MB, profiles/r08a)
                    tma_load_5d_hint(slab_s + (unsigned)sp.prow0[l] * kRowBytes, &maps.level[l], bars + 8 * l, 0, m, -1, -1, n, l2_policy_evict_first());
                }
            }
        }
        return;
    }

    // ---- consumer warps --------------------------------------------------------------------------------------------------------
    const int ncons = ncw * 32;
    constexpr int rpw = 16;  // (query, head) rows per warp
    int ql = threadIdx.x >> 1;
    const int part = threadIdx.x & 1;
    pin(ql);
    unsigned off[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        off[k] = (part * CH + ((k + ql) % CH)) * 16;
        pin(off[k]);
    }
    pin(slab_s);
    auto pass_range = [&](int u, int &nm, int &q0, int &cnt) {
        nm = u / sp.pps;
        q0 = (u - nm * sp.pps) * sp.qpp;
        cnt = min(sp.qpp, sp.Lq - q0);
        if (cnt < 0) cnt = 0;
    };
    {
        int nm, q0, cnt;
        pass_range(u_lo, nm, q0, cnt);
        stage_meta(meta0, loc, attn, sp, LP, nm, q0, cnt, fused ? refs0 : (float *)nullptr, ref, rpw);
    }
    int it = 0, resident = nm_first;
    for (int u = u_lo; u < u_hi; ++u) {
        int nm, q0, cnt;
        pass_range(u, nm, q0, cnt);
        const bool fresh = u == u_lo || nm != resident;  // first pass on this slab: its levels arrive on full[l]
        if (nm != resident) {
            resident = nm;
            ++it;
        }
        const bool release = nm != nm_last && (u + 1 - nm * sp.pps) == sp.pps;  // last pass on a slab that is followed by another
        float *meta = meta0 + ((u - u_lo) & 1) * meta_stride;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (u + 1 < u_hi) {  // prefetch the next pass's metadata while this one computes
            int nm2, q02, cnt2;
            pass_range(u + 1, nm2, q02, cnt2);
            stage_meta(meta0 + ((u + 1 - u_lo) & 1) * meta_stride, loc, attn, sp, LP, nm2, q02, cnt2,
                       fused ? refs0 + ((u + 1 - u_lo) & 1) * ref_stride : (float *)nullptr, ref, rpw);
        }
        const bool live = ql < cnt;
        const int q = q0 + ql;
        const long long n = nm / sp.M;
        const int m = nm - (int)n * sp.M;
        const long long qm = (n * sp.Lq + q) * sp.M + m;
        float *row = meta + ql * 3 * LP;
        const float2 *loc2 = reinterpret_cast<const float2 *>(row);
        const float *att = row + 2 * LP;
        if (fused) {
            // Module prologue in place on the staged row (ops/modules/ms_deform_attn.py:102-116); the two lanes of the pair run one
            // instruction stream on the two halves of the row: softmax over the L*P logits (max and sum exchanged with one shuffle
            // each; ex2-based exp, one reciprocal) and location = reference point + offset * (1/W_l, 1/H_l) (one FMA).
            const int hl = LP >> 1;  // LP % 4 == 0
            float *lg = row + 2 * LP + part * hl;
            const float *rf = refs0 + ((u - u_lo) & 1) * ref_stride + ql * sp.L * 2;
            if (LP == 12 && P == 4) {
                // the production shape (3 levels x 4 points): the lane's half row lives in registers, vector loads / stores only
                float2 g0 = make_float2(0.f, 0.f), g1 = g0, g2 = g0;
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0, o2 = o0;
                float4 *xy4 = reinterpret_cast<float4 *>(row) + part * 3;
                if (live) {
                    g0 = reinterpret_cast<const float2 *>(lg)[0]; g1 = reinterpret_cast<const float2 *>(lg)[1]; g2 = reinterpret_cast<const float2 *>(lg)[2];
                    o0 = xy4[0]; o1 = xy4[1]; o2 = xy4[2];
                }
                float mx = fmaxf(fmaxf(fmaxf(g0.x, g0.y), fmaxf(g1.x, g1.y)), fmaxf(g2.x, g2.y));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                g0.x = __expf(g0.x - mx); g0.y = __expf(g0.y - mx); g1.x = __expf(g1.x - mx);
                g1.y = __expf(g1.y - mx); g2.x = __expf(g2.x - mx); g2.y = __expf(g2.y - mx);
                float sum = (g0.x + g0.y) + (g1.x + g1.y) + (g2.x + g2.y);
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                if (live) {
                    const float inv = __frcp_rn(sum);
                    g0.x *= inv; g0.y *= inv; g1.x *= inv; g1.y *= inv; g2.x *= inv; g2.y *= inv;
                    reinterpret_cast<float2 *>(lg)[0] = g0; reinterpret_cast<float2 *>(lg)[1] = g1; reinterpret_cast<float2 *>(lg)[2] = g2;
                    float2 *ao = reinterpret_cast<float2 *>(attn_out + qm * LP + part * hl);
                    ao[0] = g0; ao[1] = g1; ao[2] = g2;
                    // samples part*6 .. part*6+5: part 0 -> levels 0,0,0,0,1,1; part 1 -> levels 1,1,2,2,2,2
                    const int la = part, lb = part + 1;
                    const float wa = sp.rw[la], ha = sp.rh[la], wb = sp.rw[lb], hb = sp.rh[lb];
                    const float xa = rf[2 * la], ya = rf[2 * la + 1], xb = rf[2 * lb], yb = rf[2 * lb + 1];
                    const bool p0 = part == 0;  // the middle float4 (samples 2, 3 of the half) belongs to level a for part 0, b for part 1
                    o0.x = __fmaf_rn(o0.x, wa, xa); o0.y = __fmaf_rn(o0.y, ha, ya); o0.z = __fmaf_rn(o0.z, wa, xa); o0.w = __fmaf_rn(o0.w, ha, ya);
                    o1.x = __fmaf_rn(o1.x, p0 ? wa : wb, p0 ? xa : xb); o1.y = __fmaf_rn(o1.y, p0 ? ha : hb, p0 ? ya : yb);
                    o1.z = __fmaf_rn(o1.z, p0 ? wa : wb, p0 ? xa : xb); o1.w = __fmaf_rn(o1.w, p0 ? ha : hb, p0 ? ya : yb);
                    o2.x = __fmaf_rn(o2.x, wb, xb); o2.y = __fmaf_rn(o2.y, hb, yb); o2.z = __fmaf_rn(o2.z, wb, xb); o2.w = __fmaf_rn(o2.w, hb, yb);
                    xy4[0] = o0; xy4[1] = o1; xy4[2] = o2;
                }
            } else {
                float mx = -3.0e38f;
                if (live)
                    for (int i = 0; i < hl; ++i) mx = fmaxf(mx, lg[i]);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                float sum = 0.f;
                if (live)
                    for (int i = 0; i < hl; ++i) {
                        const float e = __expf(lg[i] - mx);
                        lg[i] = e;
                        sum += e;
                    }
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                if (live) {
                    const float inv = __frcp_rn(sum);
                    float2 *ao = reinterpret_cast<float2 *>(attn_out + qm * LP + part * hl);
                    for (int i = 0; i < hl; i += 2) {
                        const float2 a2 = make_float2(lg[i] * inv, lg[i + 1] * inv);
                        *reinterpret_cast<float2 *>(lg + i) = a2;
                        ao[i >> 1] = a2;
                    }
                }
                if (live) {
                    float2 *xy = reinterpret_cast<float2 *>(row) + part * hl;
                    for (int i = 0; i < hl; ++i) {
                        const int l = (part * hl + i) / P;
                        float2 v = xy[i];
                        v.x = __fmaf_rn(v.x, sp.rw[l], rf[2 * l]);
                        v.y = __fmaf_rn(v.y, sp.rh[l], rf[2 * l + 1]);
                        xy[i] = v;
                    }
                }
            }
            __syncwarp();  // both lanes of a pair read the whole row below
        }
        f4p acc[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) acc[k].lo = acc[k].hi = make_float2(0.f, 0.f);
        for (int l = 0; l < sp.L; ++l) {
            if (fresh) {
                mbar_wait(bars + 8 * l, (unsigned)(it & 1));  // level l of this slab has landed (border included)
                if (fused && mask != nullptr) {  // padding mask folded in (ms_deform_attn.py:103-104): masked pixels read as zero rows
                    const int W = sp.W[l], npix = sp.H[l] * W;
                    const unsigned char *mk = mask + n * sp.S + sp.start[l];
                    float4 *lv = slab + (size_t)(sp.prow0[l] + (W + 2) + 1) * 8;
                    for (int pix = threadIdx.x; pix < npix; pix += ncons) {
                        if (mk[pix]) {
                            const int y = (int)__umulhi((unsigned)pix, sp.wrecip[l]), x = pix - y * W;
                            float4 *prow = lv + (size_t)(y * (W + 2) + x) * 8;
#pragma unroll
                            for (int c = 0; c < 8; ++c) prow[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    fence_proxy_async();  // these generic-proxy writes precede the next TMA fill of the region
                    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
                }
            }
            if (live) {
                const float Hf = (float)sp.H[l], Wf = (float)sp.W[l];
                const float hmax = sp.hmax[l], wmax = sp.wmax[l];
                const int pitch = sp.W[l] + 2;
                const unsigned lvl_s = slab_s + (unsigned)(sp.prow0[l] + pitch + 1) * kRowBytes;  // address of pixel (0,0)
                const unsigned pitch_b = (unsigned)pitch * kRowBytes;
#pragma unroll
                for (int pp = 0; pp < (P_T > 0 ? P_T : 1); ++pp) {
                    for (int p = (P_T > 0 ? pp : 0); p < (P_T > 0 ? pp + 1 : P); ++p) {
                        const int s = l * P + p;
                        const float2 xy = loc2[s];
                        const float a_in = att[s];
                        const float h = __fmaf_rn(xy.y, Hf, -0.5f), w = __fmaf_rn(xy.x, Wf, -0.5f);
                        const bool ok = h > -1.f && w > -1.f && h < Hf && w < Wf;
                        const float hc = fminf(fmaxf(h, -1.f), hmax), wc = fminf(fmaxf(w, -1.f), wmax);
                        float fh, fw;
                        int y0, x0;
                        floor_split(hc, fh, y0);
                        floor_split(wc, fw, x0);
                        const float lh = hc - fh, lw = wc - fw;
                        const unsigned r00 = lvl_s + (unsigned)(y0 * pitch + x0) * kRowBytes;
                        const unsigned r10 = r00 + pitch_b;
                        const float a = ok ? a_in : 0.f;
                        const float ahh = a * (1.f - lh), alh = a * lh, hw = 1.f - lw;
                        const float w1s = ahh * hw, w2s = ahh * lw, w3s = alh * hw, w4s = alh * lw;
                        const float2 w1 = make_float2(w1s, w1s), w2 = make_float2(w2s, w2s), w3 = make_float2(w3s, w3s), w4 = make_float2(w4s, w4s);
#pragma unroll
                        for (int k = 0; k < CH; ++k) {
                            const float4 v1 = lds4(r00 + off[k]), v2 = lds4(r00 + off[k] + kRowBytes);
                            const float4 v3 = lds4(r10 + off[k]), v4 = lds4(r10 + off[k] + kRowBytes);
                            acc[k].lo = __ffma2_rn(w1, make_float2(v1.x, v1.y), __ffma2_rn(w2, make_float2(v2.x, v2.y),
                                        __ffma2_rn(w3, make_float2(v3.x, v3.y), __ffma2_rn(w4, make_float2(v4.x, v4.y), acc[k].lo))));
                            acc[k].hi = __ffma2_rn(w1, make_float2(v1.z, v1.w), __ffma2_rn(w2, make_float2(v2.z, v2.w),
                                        __ffma2_rn(w3, make_float2(v3.z, v3.w), __ffma2_rn(w4, make_float2(v4.z, v4.w), acc[k].hi))));
                        }
                    }
                }
            }
            if (release) {  // this warp is done with level l of the resident slab: the producer may refill it
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + 8 * (kTmaMaxLevels + l));
            }
        }
        if (live) {
            float *o = out + qm * kD;
#pragma unroll
            for (int k = 0; k < CH; ++k) store4(o + (off[k] >> 2), make_float4(acc[k].lo.x, acc[k].lo.y, acc[k].hi.x, acc[k].hi.y));
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// choose passes-per-slab so that N*M*pps units spread evenly over `ctas` persistent CTAs
void plan_passes(int Lq, long long slabs, int ctas, int max_pairs, int &qpp, int &pps)
{
    const int min_pps = (Lq + max_pairs - 1) / max_pairs;
    static const int forced = [] { const char *e = getenv("MSDA_SLAB_PPS"); return e ? atoi(e) : 0; }();  // A/B switch
    if (forced >= min_pps) {
        pps = forced;
        qpp = (Lq + forced - 1) / forced;
        return;
    }
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

// plan shared by the cp.async kernel and the TMA kernel; `max_threads` bounds the lanes that gather (two per query)
bool make_params(const Problem &p, SlabParams &sp, size_t &smem_bytes, int &threads, int &grid, bool &stage, int max_threads)
{
    if (!((p.dtype == 0 || p.dtype == 1) && p.D == kD && p.have_host && p.L <= kMaxLevels)) return false;
    sp.L = p.L; sp.P = p.P; sp.S = p.S; sp.M = p.M; sp.Lq = p.Lq;
    int rows = 0;
    for (int l = 0; l < p.L; ++l) {
        sp.H[l] = p.table.H[l];
        sp.W[l] = p.table.W[l];
        sp.start[l] = p.table.start[l];
        sp.prow0[l] = rows;
        sp.hmax[l] = nextafterf((float)sp.H[l], 0.f);
        sp.wmax[l] = nextafterf((float)sp.W[l], 0.f);
        sp.rw[l] = 1.0f / (float)sp.W[l];
        sp.rh[l] = 1.0f / (float)sp.H[l];
        sp.wrecip[l] = (unsigned)((0x100000000ULL + sp.W[l] - 1) / sp.W[l]);
        if ((long long)sp.H[l] * sp.W[l] >= 65536) return false;
        rows += (sp.H[l] + 2) * (sp.W[l] + 2);
        if (rows > (1 << 20)) return false;
    }
    sp.padded_rows = rows;
    smem_bytes = (size_t)rows * kRowBytes;
    if (smem_bytes + 4096 > (size_t)p.smem_optin) return false;
    const long long slabs = (long long)p.N * p.M;
    grid = (int)(slabs * 16 < p.sm_count ? slabs * 16 : p.sm_count);
    if (grid < 1) grid = 1;
    plan_passes(p.Lq, slabs, grid, max_threads / 2, sp.qpp, sp.pps);
    sp.units = slabs * sp.pps;
    if (sp.units >= (1LL << 30)) return false;  // 32-bit unit arithmetic in the kernel
    threads = ((sp.qpp * 2 + 31) / 32) * 32;
    if (threads > max_threads) threads = max_threads;
    // metadata staging (two buffers of qpp rows of 3*L*P floats) when 16-byte pieces line up and it fits
    const int LP = p.L * p.P;
    sp.precip = LP >= 4 ? (unsigned)((0x100000000ULL + (3 * LP / 4) - 1) / (3 * LP / 4)) : 0;
    const size_t meta_bytes = 2 * (size_t)sp.qpp * 3 * LP * (p.loc_dtype == 1 ? 2 : 4);
    stage = (LP % 4 == 0) && smem_bytes + meta_bytes <= (size_t)p.smem_optin;
    if (stage) smem_bytes += meta_bytes;
    sp.wstage = stage ? 1 : 0;  // every warp stages the rows of its own lanes
    sp.fused = 0;
    if (p.ref != nullptr) {  // fused prologue: needs the staged path, float32 metadata and room for the reference points
        const size_t ref_bytes = 2 * (size_t)sp.qpp * p.L * 2 * 4;
        if (!stage || p.loc_dtype != 0 || smem_bytes + ref_bytes > (size_t)p.smem_optin) return false;
        smem_bytes += ref_bytes;
        sp.fused = 1;
    }
    return true;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is read-only state afterwards
template <auto Kern> cudaError_t ensure_smem_optin(const Problem &p)
{
    static std::atomic<unsigned long long> done{0};
    const unsigned long long bit = 1ULL << (p.device & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_optin);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

template <typename TV, typename TM>
cudaError_t launch_typed(const Problem &p, void *out, cudaStream_t st, const SlabParams &sp, size_t smem, int threads, int grid, bool stage)
{
    auto go = [&](auto kern_c) {
        constexpr auto kern = decltype(kern_c)::value;
        cudaError_t e = ensure_smem_optin<kern>(p);
        if (e != cudaSuccess) return e;
        kern<<<grid, threads, smem, st>>>((const TV *)p.value, (const TM *)p.loc, (const TM *)p.attn, (TV *)out, (const TM *)p.ref, (TM *)p.attn_out, p.mask, sp);
        count_launch();
        return cudaGetLastError();
    };
    using std::integral_constant;
    if (p.P == 4) return stage ? go(integral_constant<decltype(&fwd_slab32_kernel<TV, TM, 4, true, 2>), &fwd_slab32_kernel<TV, TM, 4, true, 2>>())
                               : go(integral_constant<decltype(&fwd_slab32_kernel<TV, TM, 4, false, 2>), &fwd_slab32_kernel<TV, TM, 4, false, 2>>());
    return stage ? go(integral_constant<decltype(&fwd_slab32_kernel<TV, TM, 0, true, 2>), &fwd_slab32_kernel<TV, TM, 0, true, 2>>())
                 : go(integral_constant<decltype(&fwd_slab32_kernel<TV, TM, 0, false, 2>), &fwd_slab32_kernel<TV, TM, 0, false, 2>>());
}

// The TMA kernel applies to float32 values + float32 metadata, L <= 4, L*P % 4 == 0, levels that fit a 256 x 256 box and a
// 16-byte aligned value pointer; everything else keeps the cp.async kernel (forward) / the sorted backward.
bool make_tma(const Problem &p, SlabParams &sp, size_t &smem, int &threads, int &grid, TmaMaps *maps)
{
    static const bool off = [] { const char *e = getenv("MSDA_SLAB_TMA"); return e && atoi(e) == 0; }();  // A/B switch
    if (off || p.dtype != 0 || p.loc_dtype != 0 || p.L > kTmaMaxLevels) return false;
    bool stage;
    if (!make_params(p, sp, smem, threads, grid, stage, kFwdTmaThreads - 32) || !stage) return false;
    if (p.ref != nullptr && p.attn_out == nullptr) return false;
    smem += 2 * kTmaMaxLevels * 8;  // full[l], empty[l]
    if (smem > (size_t)p.smem_optin) return false;
    threads += 32;  // the producer warp
    for (int l = 0; l < p.L; ++l) {
        if (sp.H[l] + 2 > 256 || sp.W[l] + 2 > 256) return false;
        if (maps != nullptr &&
            !tma_make_level_map(&maps->level[l], (const float *)p.value + (size_t)sp.start[l] * p.M * kD, p.N, p.S, p.M, sp.H[l], sp.W[l]))
            return false;
    }
    if (maps == nullptr && (tma_encode_fn() == nullptr || (reinterpret_cast<uintptr_t>(p.value) & 15) != 0)) return false;
    return true;
}

cudaError_t launch_tma(const Problem &p, void *out, cudaStream_t st, const SlabParams &sp, size_t smem, int threads, int grid, const TmaMaps &maps)
{
    auto go = [&](auto kern_c) {
        constexpr auto kern = decltype(kern_c)::value;
        cudaError_t e = ensure_smem_optin<kern>(p);
        if (e != cudaSuccess) return e;
        kern<<<grid, threads, smem, st>>>((const float *)p.loc, (const float *)p.attn, (float *)out, (const float *)p.ref, (float *)p.attn_out, p.mask, sp, maps);
        count_launch();
        return cudaGetLastError();
    };
    using std::integral_constant;
    if (p.P == 4) return go(integral_constant<decltype(&slab32_tma_kernel<4>), &slab32_tma_kernel<4>>());
    return go(integral_constant<decltype(&slab32_tma_kernel<0>), &slab32_tma_kernel<0>>());
}

}  // namespace

bool slab32_supported(const Problem &p)
{
    SlabParams sp;
    size_t smem;
    int threads, grid;
    bool stage;
    if (!make_params(p, sp, smem, threads, grid, stage, 512)) return false;
    return p.ref == nullptr || p.attn_out != nullptr;
}

cudaError_t fwd_slab32(const Problem &p, void *out, cudaStream_t st)
{
    SlabParams sp;
    size_t smem;
    int threads, grid;
    bool stage;
    if ((long long)p.N * p.M * p.Lq == 0) return cudaSuccess;
    {
        TmaMaps maps;
        if (make_tma(p, sp, smem, threads, grid, &maps)) return launch_tma(p, out, st, sp, smem, threads, grid, maps);
    }
    if (!make_params(p, sp, smem, threads, grid, stage, 512)) return cudaErrorInvalidValue;
    if (p.ref != nullptr && p.attn_out == nullptr) return cudaErrorInvalidValue;
    return p.dtype == 0                    ? launch_typed<float, float>(p, out, st, sp, smem, threads, grid, stage)
           : p.loc_dtype == 0              ? launch_typed<bf16, float>(p, out, st, sp, smem, threads, grid, stage)
                                           : launch_typed<bf16, bf16>(p, out, st, sp, smem, threads, grid, stage);
}

}  // namespace msda
