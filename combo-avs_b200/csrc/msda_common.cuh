// msda_common.cuh -- shared device helpers for the MSDeformAttn kernels (sm_100a only).
//
// Arithmetic contract (reference: ops/src/cuda/ms_deform_im2col_cuda.cuh:38-89, 277-302):
//   h = loc_y*H - 0.5, w = loc_x*W - 0.5 (one FMA, as the reference's compiled kernel does)
//   sample contributes iff h > -1 && w > -1 && h < H && w < W
//   neighbours (y0,x0) (y0,x1) (y1,x0) (y1,x1), y0 = floor(h); a neighbour outside the map reads as zero
//   weights hh*hw, hh*lw, lh*hw, lh*lw with lh = h - y0, hh = 1 - lh.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

constexpr int kMaxLevels = 16;  // geometry table size in shared / constant parameter blocks

// Level geometry handed to kernels by value when the host knows it (shapes_host != NULL).
struct LevelTable {
    int H[kMaxLevels];
    int W[kMaxLevels];
    int start[kMaxLevels];
};

template <typename T> struct Compute { using type = float; };
template <> struct Compute<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename Compute<T>::type to_compute(T v) { return v; }
template <> __device__ __forceinline__ float to_compute<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T, typename CT> __device__ __forceinline__ T from_compute(CT v) { return (T)v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_compute<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float pix_coord(float loc, int n) { return __fmaf_rn(loc, (float)n, -0.5f); }
__device__ __forceinline__ double pix_coord(double loc, int n) { return fma(loc, (double)n, -0.5); }

// Four-neighbour footprint of one sample.
template <typename CT> struct Taps {
    int y0, x0;          // top-left neighbour (may be -1)
    CT lh, lw, hh, hw;   // fractional parts and complements
    bool top, bot, lef, rig;  // neighbour rows / columns that lie inside the map
};

template <typename CT>
__device__ __forceinline__ bool make_taps(CT loc_x, CT loc_y, int H, int W, Taps<CT> &t)
{
    const CT h = pix_coord(loc_y, H);
    const CT w = pix_coord(loc_x, W);
    if (!(h > (CT)-1 && w > (CT)-1 && h < (CT)H && w < (CT)W)) return false;
    const CT fh = floor(h), fw = floor(w);
    t.y0 = (int)fh;
    t.x0 = (int)fw;
    t.lh = h - fh;
    t.lw = w - fw;
    t.hh = (CT)1 - t.lh;
    t.hw = (CT)1 - t.lw;
    t.top = t.y0 >= 0;
    t.bot = t.y0 + 1 <= H - 1;
    t.lef = t.x0 >= 0;
    t.rig = t.x0 + 1 <= W - 1;
    return true;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// 16-byte vector reduction to global memory (sm_90+): one L2 atomic transaction for four floats.
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// the same, predicated inside the instruction: no branch / reconvergence around it when only some lanes reduce
__device__ __forceinline__ void red_add_v4_if(int pred, float *p, float a, float b, float c, float d)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "r"(pred) : "memory");
}

__device__ __forceinline__ void atomic_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(double *p, double v) { atomicAdd(p, v); }

template <typename CT> __device__ __forceinline__ CT warp_sum(CT v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace msda
