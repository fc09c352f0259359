// msda_linear.cu -- row f2 of the scope table: the Linear layers AROUND the op in the deformable encoder layer
// (reference models/modeling/pixel_decoder/msdeformattn.py:119-134: linear1 256->1024, linear2 1024->256;
//  ops/modules/ms_deform_attn.py:105-128: value_proj / sampling_offsets / attention_weights / output_proj) on the 5th-generation
// tensor cores, at float32 accuracy.
//
// The reference runs these GEMMs in strict fp32 (AMP is off for S4 / MS3, matmul.allow_tf32 False): 1.1 TFLOP per encoder training step
// on the SIMT pipes, 85 % of the step.  Here every float32 operand is split into two TF32-representable halves,
// x = hi + lo with hi = x with its 13 low mantissa bits cleared and lo = x - hi (exact), and
//     x * w  ~=  hi_x*hi_w + hi_x*lo_w + lo_x*hi_w            (the dropped lo*lo term is below 2^-21 relative)
// is accumulated in float32 in tensor memory by three tcgen05.mma.kind::tf32 per 8-wide K step ("3xTF32").
//
// Kernel (one persistent CTA per SM, warp-specialised, 448 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled boxes of A [128 x 32] and W [BN x 32] into a ring of stages
//   warps 2-9   splitters: when a stage has landed, write lo = x - trunc_tf32(x) next to it (shared memory only; kind::tf32 ignores the
//               13 low mantissa bits of its operands -- measured -- so the raw tile already is `hi`), fence.proxy.async, hand the stage to
//   warp 1      the MMA issuer: one elected lane issues 3 x 4 tcgen05.mma (M=128, N=BN<=128, K=8) per stage -- hi*hi into one TMEM
//               accumulator, hi*lo + lo*hi into a second one (the tensor core truncates when it aligns addends; kept apart, the small
//               cross terms are truncated at their own magnitude and the large accumulator sees a third of the additions);
//               tcgen05.commit releases the stage to the producer and, after the last K block, the accumulators to
//   warps 10-13 the epilogue: tcgen05.ld both accumulators (lane = row), add them, the bias (from shared memory) and optionally ReLU or a
//               ReLU' mask, transpose 32 x 32 blocks through a padded staging tile and write 128 contiguous bytes per row (a direct
//               row store touches 32 lines with 16 bytes each: measured 143 us for the stores alone); the next tile's accumulators
//               (the other 256 TMEM columns) are being filled meanwhile.
// C[M, N] = A[M, K] . W[N, K]^T + bias[N] serves the forward (W = weight) and dgrad (A = grad_y, W = weight^T).  The weight gradient
// C[N, K] = grad_y[M, N]^T . x[M, K] contracts over the ROWS of both operands: the same kernel in MN-major mode (32 x 32 TMA boxes,
// transposed-operand bits in the instruction descriptor) reads grad_y and x as they lie in memory, and splits the long contraction over
// the SMs (partial tiles are added into the zeroed result with red.global.add.v4.f32).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "../../include/msda_b200.h"
#include "msda_kernels.h"
#include "msda_tma.cuh"

namespace msda {
namespace {

constexpr int kBM = 128;          // rows of A per tile = TMEM lanes
constexpr int kBK = 32;           // floats per K block = one 128-byte swizzle row
constexpr int kUmmaK = 8;         // K of one tcgen05.mma.kind::tf32
constexpr int kSplitters = 8;  // splitter warps (4 / 8 / 12 / 16: 164 / 137 / 140 / 141 us on 41160 x 256 x 1024, profiles/r08vw_linear_splitters.txt)
constexpr int kSplitWarp0 = 2, kEpiWarp0 = kSplitWarp0 + kSplitters;
constexpr int kThreads = (kEpiWarp0 + 4) * 32;
constexpr int kStagePitch = 36;   // floats per staged row: 32 columns + 4 of padding (16-byte aligned, conflict-free both ways)
constexpr int kBiasFloats = 2048;

struct LinearParams {
    int M, N, K;
    int BN;            // columns of C per tile (multiple of 16, <= 256, divides N)
    int stages;
    int split;         // 1: 3xTF32 (float32 accuracy); 0: plain TF32 (profiling / bring-up only)
    int relu;
    int mn_major;      // 1: C[N, K] = A[Mr, N]^T . W[Mr, K] -- both operands are read transposed (MN-major), the contraction runs over their rows
    int ksplit;        // > 1: the K blocks of a tile are shared out over `ksplit` work items, which ADD their partial sums into a zeroed C
    int kb_total, kb_per;
    int tiles_m, tiles_n;
    unsigned stage_bytes, a_bytes, b_bytes;
};

struct alignas(64) LinearMaps {
    CUtensorMap a, w;
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 2-D box of a tiled tensor map -> shared memory
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// shared-memory matrix descriptor of a K-major tile whose rows are one 128-byte swizzle span (cute::UMMA::SmemDescriptor):
// start address >> 4 in bits [0,14), stride between 8-row groups (1024 B) >> 4 in bits [32,46), version 1 in bits [46,48),
// layout SWIZZLE_128B (2) in bits [61,64); the leading-dimension offset is not used by this layout
__device__ __forceinline__ unsigned long long umma_desc(unsigned smem_addr)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3ffff) >> 4);
    d |= (unsigned long long)(1024 >> 4) << 32;
    d |= 1ULL << 46;
    d |= 2ULL << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
// (bit 15 / 16: A / B is MN-major, i.e. stored transposed)
__device__ __forceinline__ unsigned umma_idesc(int m, int n, int mn_major)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(mn_major ? 3 : 0) << 15) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

// MN-major operand tile made of 32 (K) x 32 (MN) float boxes, 4096 bytes each, one after the other along MN.  For 32-bit MN-major
// operands the tensor core accepts exactly one shared-memory layout (cutlass sm100_common.inl:92): the 128-byte swizzle with a 32-byte
// base -- TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, descriptor layout type 1, cute Swizzle<2,5,2> o ((8,n),(4,k)):((1,LBO),(8,SBO)) in
// 16-byte units: an atom is 4 K-rows x 32 MN-floats = 512 bytes.  So LBO = 4096 bytes between MN atoms (the next box), SBO = 512 bytes
// between K atoms (the next four rows of the box), and one K = 8 step advances the start address by 1024 bytes.
__device__ __forceinline__ unsigned long long umma_desc_mn(unsigned smem_addr)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr & 0x3ffff) >> 4);
    d |= (unsigned long long)(4096 >> 4) << 16;
    d |= (unsigned long long)(512 >> 4) << 32;
    d |= 1ULL << 46;
    d |= 1ULL << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// all tcgen05.mma issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16])
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kThreads, 1)
linear_tf32_kernel(const float *__restrict__ bias, float *__restrict__ C, const __grid_constant__ LinearParams lp,
                   const __grid_constant__ LinearMaps maps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    // stage s: [A hi 16 KB][W hi b_bytes] (and, with split: [A lo][W lo]) -- every tile 1024-byte aligned for the 128-byte swizzle
    // (the 128-byte swizzle wants 1024-byte aligned tiles; the launch reserves the slack for rounding the base up)
    const unsigned smem_raw = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned pad = (1024u - (smem_raw & 1023u)) & 1023u;
    unsigned char *smem_al = smem + pad;
    const unsigned smem_s = smem_raw + pad;
    const unsigned bars = smem_s + lp.stages * lp.stage_bytes;   // full[s], ready[s], empty[s], tmem_full[2], tmem_empty[2], tmem base
    auto full = [&](int s) { return bars + 8 * s; };
    auto ready = [&](int s) { return bars + 8 * (lp.stages + s); };
    auto empty = [&](int s) { return bars + 8 * (2 * lp.stages + s); };
    auto tfull = [&](int b) { return bars + 8 * (3 * lp.stages + b); };
    auto tempty = [&](int b) { return bars + 8 * (3 * lp.stages + 2 + b); };
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem_al + lp.stages * lp.stage_bytes + 8 * (3 * lp.stages + 4));
    float *stage0 = reinterpret_cast<float *>(smem_al + lp.stages * lp.stage_bytes + 256);   // 4 epilogue warps x 32 rows x kStagePitch floats
    float *sbias = stage0 + 4 * 32 * kStagePitch;                                            // the bias vector, when it fits kBiasFloats
    const bool have_sbias = bias != nullptr && lp.N <= kBiasFloats;
    if (have_sbias)
        for (int i = threadIdx.x; i < lp.N; i += kThreads) sbias[i] = __ldg(bias + i);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = lp.tiles_m * lp.tiles_n * lp.ksplit;
    // work item t -> (tile row, tile column, K slice): its K blocks are [kb0, kb0 + kblocks)
    auto decode = [&](int t, int &tm, int &tn, int &kb0, int &kblocks) {
        const int ks = t % lp.ksplit;
        const int tt = t / lp.ksplit;
        tm = tt / lp.tiles_n;
        tn = tt - tm * lp.tiles_n;
        kb0 = ks * lp.kb_per;
        kblocks = min(lp.kb_per, lp.kb_total - kb0);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < lp.stages; ++s) {
            mbar_init(full(s), 1);
            mbar_init(ready(s), kSplitters);   // one arrival per splitter warp
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull(b), 1);
            mbar_init(tempty(b), 4);  // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (wid == 1) {  // the MMA warp owns the tensor-memory allocation: 512 columns = two float32 accumulators of up to 256 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (wid == 0) {
        // ---- TMA producer ---------------------------------------------------------------------------------------------------
        if (lane == 0) {
            tma_prefetch_map(&maps.a);
            tma_prefetch_map(&maps.w);
            int it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int tm, tn, kb0, kblocks;
                decode(t, tm, tn, kb0, kblocks);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % lp.stages, round = it / lp.stages;
                    if (round > 0) mbar_wait(empty(s), (unsigned)((round - 1) & 1));
                    const unsigned base = smem_s + s * lp.stage_bytes;
                    mbar_arrive_expect_tx(full(s), lp.a_bytes + lp.b_bytes);
                    if (lp.mn_major) {  // 32 x 32 boxes: 32 rows of the contraction, 32 columns of the operand (rows / columns past the end arrive as zeros)
                        for (int j = 0; j < kBM / 32; ++j) tma_load_2d(base + j * 4096, &maps.a, full(s), tm * kBM + j * 32, (kb0 + kb) * kBK);
                        for (int j = 0; j < lp.BN / 32; ++j) tma_load_2d(base + lp.a_bytes + j * 4096, &maps.w, full(s), tn * lp.BN + j * 32, (kb0 + kb) * kBK);
                    } else {
                        tma_load_2d(base, &maps.a, full(s), (kb0 + kb) * kBK, tm * kBM);       // (columns past K arrive as zeros)
                        tma_load_2d(base + lp.a_bytes, &maps.w, full(s), (kb0 + kb) * kBK, tn * lp.BN);
                    }
                }
            }
        }
    } else if (wid == 1) {
        // ---- MMA issuer -----------------------------------------------------------------------------------------------------
        const unsigned idesc = umma_idesc(kBM, lp.BN, lp.mn_major);
        int it = 0, lt = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++lt) {
            int tm, tn, kb0, kblocks;
            decode(t, tm, tn, kb0, kblocks);
            const int buf = lt & 1, use = lt >> 1;
            if (use > 0) mbar_wait(tempty(buf), (unsigned)((use - 1) & 1));  // the epilogue has drained this accumulator
            tc_fence_after();
            // split mode keeps TWO accumulators per tile: the hi*hi products in one, the 2^-11 times smaller cross terms in the other.
            // The tensor core truncates when it aligns addends to the accumulator, an error proportional to the accumulator's
            // magnitude and always in the same direction; kept apart, the cross terms are truncated at THEIR magnitude (negligible)
            // and the large accumulator sees a third of the additions.  The epilogue adds the two in float32.
            // (two tiles in flight, 256 TMEM columns each: plain TF32 one accumulator of BN <= 256 columns; split: BN <= 128, hi*hi in
            // columns 0-127 and the cross terms in 128-255)
            const unsigned tmem_d = tmem_base + buf * 256, tmem_x = tmem_d + 128;
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const int s = it % lp.stages, round = it / lp.stages;
                mbar_wait(lp.split ? ready(s) : full(s), (unsigned)(round & 1));
                tc_fence_after();
                if (lane == 0) {
                    const unsigned a_hi = smem_s + s * lp.stage_bytes, w_hi = a_hi + lp.a_bytes;
                    const unsigned a_lo = w_hi + lp.b_bytes, w_lo = a_lo + lp.a_bytes;
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        // K-major: 32 bytes along the swizzled 128-byte row; MN-major: the next 8-row swizzle atom
                        const unsigned koff = lp.mn_major ? k * 1024 : k * kUmmaK * 4;
                        const unsigned long long dah = lp.mn_major ? umma_desc_mn(a_hi + koff) : umma_desc(a_hi + koff);
                        const unsigned long long dwh = lp.mn_major ? umma_desc_mn(w_hi + koff) : umma_desc(w_hi + koff);
                        umma_tf32(tmem_d, dah, dwh, idesc, (kb | k) != 0);
                        if (lp.split) {
                            const unsigned long long dal = lp.mn_major ? umma_desc_mn(a_lo + koff) : umma_desc(a_lo + koff);
                            const unsigned long long dwl = lp.mn_major ? umma_desc_mn(w_lo + koff) : umma_desc(w_lo + koff);
                            umma_tf32(tmem_x, dah, dwl, idesc, (kb | k) != 0);
                            umma_tf32(tmem_x, dal, dwh, idesc, 1);
                        }
                    }
                    umma_commit(empty(s));                       // the stage is free once these MMAs have read it
                    if (kb == kblocks - 1) umma_commit(tfull(buf));  // ... and after the last K block the accumulator is complete
                }
                __syncwarp();
            }
        }
    } else if (wid < kEpiWarp0) {
        // ---- splitters: x -> (hi, lo) in shared memory --------------------------------------------------------------------------
        if (lp.split) {
            const int st = threadIdx.x - kSplitWarp0 * 32;  // 0..32*kSplitters-1
            int it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int tm, tn, kb0, kblocks;
                decode(t, tm, tn, kb0, kblocks);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % lp.stages, round = it / lp.stages;
                    mbar_wait(full(s), (unsigned)(round & 1));
                    unsigned char *base = smem_al + (size_t)s * lp.stage_bytes;
                    const unsigned hi_bytes = lp.a_bytes + lp.b_bytes;  // [A | W] is contiguous, and so is [A lo | W lo] behind it
                    float4 *hi = reinterpret_cast<float4 *>(base);
                    float4 *lo = reinterpret_cast<float4 *>(base + hi_bytes);
                    const unsigned n16 = hi_bytes / 16;  // four 16-byte chunks per thread per round: the loads of a round go out together
                    for (unsigned i0 = st; i0 < n16; i0 += 4 * (kSplitters * 32)) {
                        float4 x[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (i0 + r * (kSplitters * 32) < n16) x[r] = hi[i0 + r * (kSplitters * 32)];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (i0 + r * (kSplitters * 32) >= n16) break;
                            // kind::tf32 ignores the 13 low mantissa bits of its operands (measured, profiles/r05_linear_tcgen05.txt): the raw
                            // float32 tile already IS the hi operand; only lo = x - trunc(x) (exact in float32) is written
                            float4 l;
                            l.x = x[r].x - __uint_as_float(__float_as_uint(x[r].x) & 0xffffe000u);
                            l.y = x[r].y - __uint_as_float(__float_as_uint(x[r].y) & 0xffffe000u);
                            l.z = x[r].z - __uint_as_float(__float_as_uint(x[r].z) & 0xffffe000u);
                            l.w = x[r].w - __uint_as_float(__float_as_uint(x[r].w) & 0xffffe000u);
                            lo[i0 + r * (kSplitters * 32)] = l;
                        }
                    }
                    fence_proxy_async();  // these generic-proxy writes are read by the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ready(s));
                }
            }
        }
    } else {
        // ---- epilogue: TMEM -> registers (+ bias, ReLU) -> per-warp staging tile in shared memory -> coalesced global stores ------------
        // tcgen05.ld hands every lane one ROW of the accumulator; written straight out, a warp store would touch 32 different 128-byte
        // lines with 16 bytes each (measured: the stores alone then take 143 us for the 41160 x 1024 result, 3.6x the whole rest of the
        // kernel).  So each warp transposes 32 rows x 32 columns through a padded staging tile and writes 128 contiguous bytes per row.
        const int q = wid & 3;                        // a warp reads the TMEM lanes 32*(warp id % 4) ...
        float *stage = stage0 + (wid - kEpiWarp0) * (32 * kStagePitch);
        int lt = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++lt) {
            int tm, tn, kb0, kblocks;
            decode(t, tm, tn, kb0, kblocks);
            const int buf = lt & 1, use = lt >> 1;
            mbar_wait(tfull(buf), (unsigned)(use & 1));
            tc_fence_after();
            const int row0 = tm * kBM + q * 32;      // first row of this warp's 32
            const bool add_bias = bias != nullptr && kb0 == 0;  // (split K: the first slice carries the bias)
            const unsigned taddr = tmem_base + buf * 256 + ((unsigned)(q * 32) << 16);
            for (int c0 = 0; c0 < lp.BN;) {
                const int cw = (lp.BN - c0) >= 32 ? 32 : 16;   // columns in this round
                for (int h = 0; h < cw; h += 16) {
                    float v[16];
                    tmem_ld16(taddr + c0 + h, v);
                    if (lp.split) {
                        float x[16];
                        tmem_ld16(taddr + 128 + c0 + h, x);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += x[j];
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (add_bias) {
                            const float4 bb = have_sbias ? *reinterpret_cast<const float4 *>(sbias + tn * lp.BN + c0 + h + j)
                                                         : __ldg(reinterpret_cast<const float4 *>(bias + tn * lp.BN + c0 + h + j));
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        if (lp.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4 *>(stage + lane * kStagePitch + h + j) = o;
                    }
                }
                __syncwarp();
                const int lpr = cw >> 2, rpi = 32 / lpr;         // lanes per row, rows per store instruction
                const int rr = lane / lpr, cc = lane - rr * lpr;
                {
                    for (int r = rr; r < 32; r += rpi) {
                        const float4 o = *reinterpret_cast<const float4 *>(stage + r * kStagePitch + cc * 4);
                        const int row = row0 + r;
                        if (row < lp.M) {
                            float *dst = C + (size_t)row * lp.N + tn * lp.BN + c0 + cc * 4;
                            if (lp.ksplit > 1) red_add_v4(dst, o.x, o.y, o.z, o.w);
                            else *reinterpret_cast<float4 *>(dst) = o;
                        }
                    }
                }
                __syncwarp();
                c0 += cw;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

bool make_map_2d(CUtensorMap *map, const float *base, int rows, int cols, int box_rows, bool mn_major = false)
{
    if (box_rows > 256) return false;
    EncodeTiledFn enc = tma_encode_fn();
    if (enc == nullptr || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_bn(int N, int limit, int step = 16)
{
    for (int bn = limit; bn >= step; bn -= step)
        if (N % bn == 0) return bn;
    return 0;
}

}  // namespace

// weight-gradient form: C[N, K] = G[M, N]^T . X[M, K] (G, X row-major): N % 4 (row pitch of G), K % 32 (32-column boxes of X)
bool linear_wgrad_supported(int M, int N, int K)
{
    return M > 0 && N > 0 && K > 0 && N % 4 == 0 && K % 32 == 0 && tma_encode_fn() != nullptr;
}

bool linear_supported(int M, int N, int K)
{
    // K % 4: the row pitch of x / weight must be a multiple of 16 bytes for the tensor maps (the K tail is zero-filled by TMA)
    return M > 0 && N > 0 && K > 0 && K % 4 == 0 && N % 16 == 0 && pick_bn(N, 128) >= 16 && tma_encode_fn() != nullptr;
}

// mn_major = 0: C[M, N] = A[M, K] . W[N, K]^T + bias.   mn_major = 1: C[M, N] = A[K, M]^T . W[K, N] (the weight gradient: A = grad_y,
// W = x, K = their common row count), no bias.
cudaError_t linear_tf32(const float *A, const float *W, const float *bias, float *C, int M, int N, int K, int relu, int split, int mn_major,
                        int device, int sm_count, int smem_optin, cudaStream_t st)
{
    LinearParams lp;
    lp.M = M; lp.N = N; lp.K = K;
    lp.split = split ? 1 : 0;
    lp.mn_major = mn_major ? 1 : 0;
    // split: BN <= 128 so that a tile's two accumulators take 256 TMEM columns and two tiles are in flight (BN = 256 with one tile in
    // flight measured slower: 184 vs 167 us on 41160 x 1024 x 256, profiles/r05_linear_tcgen05.txt)
    lp.BN = lp.mn_major ? pick_bn(N, 128, 32) : pick_bn(N, lp.split ? 128 : 256);
    if (lp.BN < 16) return cudaErrorInvalidValue;
    lp.relu = relu ? 1 : 0;
    lp.tiles_m = (M + kBM - 1) / kBM;
    lp.tiles_n = N / lp.BN;
    lp.kb_total = (K + kBK - 1) / kBK;
    // split K when the output alone cannot fill the GPU (the weight gradient: a few tiles, tens of thousands of K): C is zeroed
    // first and every slice adds its partial tile with red.global.add.v4.f32
    lp.ksplit = 1;
    if (!relu && lp.tiles_m * lp.tiles_n * 2 <= sm_count && lp.kb_total >= 16) {
        lp.ksplit = sm_count / (lp.tiles_m * lp.tiles_n);
        if (lp.ksplit > lp.kb_total / 4) lp.ksplit = lp.kb_total / 4;
        if (lp.ksplit < 1) lp.ksplit = 1;
    }
    lp.kb_per = (lp.kb_total + lp.ksplit - 1) / lp.ksplit;
    lp.ksplit = (lp.kb_total + lp.kb_per - 1) / lp.kb_per;  // no empty slices
    lp.a_bytes = kBM * kBK * 4;
    lp.b_bytes = (unsigned)lp.BN * kBK * 4;
    lp.stage_bytes = (lp.a_bytes + lp.b_bytes) * (lp.split ? 2 : 1);
    // barriers + the TMEM base slot (256) + epilogue staging + bias table + slack to align the tiles to 1024 bytes
    const size_t tail = 256 + 4 * 32 * kStagePitch * 4 + kBiasFloats * 4 + 1024;
    lp.stages = (int)(((size_t)smem_optin - tail) / lp.stage_bytes);
    if (lp.stages > 8) lp.stages = 8;
    if (lp.stages < 2) return cudaErrorInvalidValue;
    LinearMaps maps;
    if (lp.mn_major) {  // operands as they lie in memory: [K rows, M or N columns], fetched in 32 x 32 boxes
        if (!make_map_2d(&maps.a, A, K, M, 32, true) || !make_map_2d(&maps.w, W, K, N, 32, true)) return cudaErrorInvalidValue;
    } else if (!make_map_2d(&maps.a, A, M, K, kBM) || !make_map_2d(&maps.w, W, N, K, lp.BN)) {
        return cudaErrorInvalidValue;
    }
    static std::atomic<unsigned long long> done{0};
    const unsigned long long bit = 1ULL << (device & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
        if (e != cudaSuccess) return e;
        done.fetch_or(bit, std::memory_order_release);
    }
    const size_t smem = (size_t)lp.stages * lp.stage_bytes + tail;
    const int tiles = lp.tiles_m * lp.tiles_n * lp.ksplit;
    const int grid = tiles < sm_count ? tiles : sm_count;
    if (lp.ksplit > 1) {
        cudaError_t e = cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st);
        if (e != cudaSuccess) return e;
    }
    linear_tf32_kernel<<<grid, kThreads, smem, st>>>(bias, C, lp, maps);
    count_launch();
    return cudaGetLastError();
}

}  // namespace msda
