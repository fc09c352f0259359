// msda_tma.cuh -- TMA (cp.async.bulk.tensor) + mbarrier helpers for the slab kernels (sm_100a).
//
// One (frame, head, level) slab of value[N, S, M, 32] is a strided 3-D box -- 128 contiguous bytes per pixel at a pitch of
// M*128 bytes, W pixels per row, H rows -- which is exactly what the tensor-memory accelerator moves with ONE instruction issued
// by ONE thread.  The box is requested one pixel larger all round than the level, starting at pixel (-1, -1): the TMA unit fills
// the out-of-bounds part with zeros, which is the zero border that removes the corner predicates of
// ms_deform_im2col_cuda.cuh:58-82 from the gather.  Completion is signalled on an mbarrier (complete_tx::bytes), so no warp waits
// on the copy it issued and nobody executes address arithmetic for it.
//
// The tensor maps are built on the host per call (the level base pointers change with the value tensor) through the driver entry
// point cuTensorMapEncodeTiled, resolved at run time with cudaGetDriverEntryPoint: the library links against cudart only.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NULL when the driver does not export the symbol (then the callers keep their cp.async path)
inline EncodeTiledFn tma_encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Tensor map of one level of value[N, S, M, D=32] (float32): dims (fastest first) {32, M, W, H, N}, box {32, 1, W+2, H+2, 1}.
// `level_base` = value + start_l * M * 32.  Returns false when the geometry does not fit a TMA box or the pointer is misaligned.
inline bool tma_make_level_map(CUtensorMap *map, const float *level_base, int N, int S, int M, int H, int W)
{
    EncodeTiledFn enc = tma_encode_fn();
    if (enc == nullptr) return false;
    if ((reinterpret_cast<uintptr_t>(level_base) & 15) != 0 || W + 2 > 256 || H + 2 > 256 || N < 1) return false;
    const cuuint64_t dims[5] = {32, (cuuint64_t)M, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[4] = {128, (cuuint64_t)M * 128, (cuuint64_t)W * M * 128, (cuuint64_t)S * M * 128};  // bytes, dims 1..4
    const cuuint32_t box[5] = {32, 1, (cuuint32_t)(W + 2), (cuuint32_t)(H + 2), 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < 4; ++i)
        if (strides[i] >= (1ULL << 40)) return false;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(level_base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one tensor map per level, handed to the kernels by value (__grid_constant__)
constexpr int kTmaMaxLevels = 4;
struct alignas(64) TmaMaps {
    CUtensorMap level[kTmaMaxLevels];
};

#ifdef __CUDACC__
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// make mbarrier initialisation visible to the async proxy (the TMA unit)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order earlier generic-proxy accesses of shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// one 5-D box of a tiled tensor map -> shared memory; completion (bytes) is counted on `bar`
__device__ __forceinline__ void tma_load_5d(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// the same with an L2 eviction policy (createpolicy) for the global-memory side of the copy
__device__ __forceinline__ void tma_load_5d_hint(unsigned dst, const CUtensorMap *map, unsigned bar, int c0, int c1, int c2, int c3, int c4,
                                                 unsigned long long policy)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(policy) : "memory");
}
// L2 eviction policies: evict_first for data that is read once (a value slab on its way into shared memory), evict_last for data
// that is re-read from L2 while a unit is processed (grad_output rows)
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace msda
