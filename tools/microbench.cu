// microbench.cu -- pipe-throughput probes that decide the MSDeformAttn kernel designs (DESIGN.md "evidence").
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench tools/microbench.cu
// Each probe mimics one candidate access pattern of the S4 shape: rows of 128 B (32 fp32 channels), random row
// indices, 1029-row slabs (131.7 KB) or the whole 42 MB value tensor.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

constexpr int ROWS = 1029;      // slab rows
constexpr int ITERS = 2048;     // row-gathers per lane-group

// (a1) smem gather, mapping A: 8 lanes x float4 = one row; 4 rows per warp instruction
__global__ void __launch_bounds__(1024, 1) lds_groups(float *out, int iters)
{
    extern __shared__ float4 slab4[];
    for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) slab4[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) / 8 * 2654435761u + 12345u;
    const int j = threadIdx.x & 7;
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const int row = lcg(s) % ROWS;
        const float4 v = slab4[row * 8 + j];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 1.2345f) out[0] = acc.x;
}

// (a2) smem gather, mapping B: every lane its own row, chunk rotated by lane so that a quarter-warp covers all banks
__global__ void __launch_bounds__(1024, 1) lds_lanes(float *out, int iters)
{
    extern __shared__ float4 slab4[];
    for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) slab4[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    const int j = threadIdx.x & 7;
    float4 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float4(0, 0, 0, 0);
    for (int i = 0; i < iters / 8; ++i) {
        const int row = lcg(s) % ROWS;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 v = slab4[row * 8 + ((k + j) & 7)];
            acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
        }
    }
    float t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += acc[k].x + acc[k].y + acc[k].z + acc[k].w;
    if (t == 1.2345f) out[0] = t;
}

// (b/c) global gather, mapping A; rows_span = 1029 (per-CTA slab: L1-resident) or all rows (L2-resident)
__global__ void __launch_bounds__(1024, 1) ldg_groups(const float4 *__restrict__ v, float *out, int iters, int rows_span, int per_cta)
{
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) / 8 * 2654435761u + 12345u;
    const int j = threadIdx.x & 7;
    const float4 *base = v + (per_cta ? (size_t)blockIdx.x * rows_span * 8 : 0);
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const int row = lcg(s) % rows_span;
        const float4 x = __ldg(base + (size_t)row * 8 + j);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 1.2345f) out[0] = acc.x;
}

// (d1) REDG.v4 scatter, mapping A: 8 lanes cover one 128 B row, 4 rows per warp instruction
__global__ void __launch_bounds__(1024, 1) red_v4_groups(float *g, int iters, int rows_span)
{
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) / 8 * 2654435761u + 777u;
    const int j = threadIdx.x & 7;
    for (int i = 0; i < iters; ++i) {
        const int row = lcg(s) % rows_span;
        float *p = g + ((size_t)row * 8 + j) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    }
}

// (d2) REDG.v4 scatter, mapping B: every lane its own row (16 B pieces), 8 instructions per row
__global__ void __launch_bounds__(1024, 1) red_v4_lanes(float *g, int iters, int rows_span)
{
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 777u;
    const int j = threadIdx.x & 7;
    for (int i = 0; i < iters / 8; ++i) {
        const int row = lcg(s) % rows_span;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float *p = g + ((size_t)row * 8 + ((k + j) & 7)) * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
        }
    }
}

// (e) scalar RED, a warp covers one 128 B row (the reference's pattern: one atomicAdd per channel)
__global__ void __launch_bounds__(1024, 1) red_scalar_rows(float *g, int iters, int rows_span)
{
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) / 32 * 2654435761u + 777u;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < iters / 4; ++i) {
        const int row = lcg(s) % rows_span;
        atomicAdd(g + (size_t)row * 32 + lane, 1.0f);
    }
}

// (f1) shared int atomics (native ATOMS.ADD), lane-distinct banks, random rows
__global__ void __launch_bounds__(1024, 1) atoms_int(int *out, int iters)
{
    extern __shared__ int sl[];
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) sl[i] = 0;
    __syncthreads();
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) / 32 * 2654435761u + 99u;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < iters / 4; ++i) {
        const int row = lcg(s) % ROWS;
        atomicAdd(&sl[row * 32 + lane], 1);
    }
    __syncthreads();
    if (sl[threadIdx.x] == -1) out[0] = 1;
}

// (f2) shared float atomics (CAS loop), lane-distinct banks, random rows
__global__ void __launch_bounds__(1024, 1) atoms_f32(float *out, int iters)
{
    extern __shared__ float sf[];
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) sf[i] = 0;
    __syncthreads();
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) / 32 * 2654435761u + 99u;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < iters / 4; ++i) {
        const int row = lcg(s) % ROWS;
        atomicAdd(&sf[row * 32 + lane], 1.0f);
    }
    __syncthreads();
    if (sf[threadIdx.x] == -1.f) out[0] = 1;
}

// (f3) shared plain read-modify-write of a row (the no-atomics owner-computes pattern): LDS.128 + STS.128 per lane
__global__ void __launch_bounds__(1024, 1) lds_sts_rmw(float *out, int iters)
{
    extern __shared__ float4 slab4[];
    for (int i = threadIdx.x; i < ROWS * 8; i += blockDim.x) slab4[i] = make_float4(0, 0, 0, 0);
    __syncthreads();
    unsigned s = (blockIdx.x * 1024 + threadIdx.x) / 8 * 2654435761u + 5u;
    const int j = threadIdx.x & 7;
    for (int i = 0; i < iters; ++i) {
        const int row = lcg(s) % ROWS;
        float4 v = slab4[row * 8 + j];
        v.x += 1.f; v.y += 2.f; v.z += 3.f; v.w += 4.f;
        slab4[row * 8 + j] = v;
    }
    __syncthreads();
    if (slab4[threadIdx.x].x == -1.f) out[0] = 1;
}

// (g) shuffle throughput, alone
__global__ void __launch_bounds__(1024, 1) shfl_only(float *out, int iters)
{
    float a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    for (int i = 0; i < iters; ++i) {
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        b += __shfl_xor_sync(0xffffffffu, b, 2);
        c += __shfl_xor_sync(0xffffffffu, c, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 8);
    }
    if (a + b + c + d == 1.2345f) out[0] = a;
}

// (h) FFMA issue rate reference: 16 independent FMAs per iteration
__global__ void __launch_bounds__(1024, 1) ffma_only(float *out, int iters, float x)
{
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], x, a[(k + 1) & 15]);
    }
    float t = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += a[k];
    if (t == 1.2345f) out[0] = t;
}

// (h2) packed FFMA2 (fma.rn.f32x2) issue rate: 16 independent packed FMAs per iteration
__global__ void __launch_bounds__(1024, 1) ffma2_only(float *out, int iters, float x)
{
    float2 a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = make_float2(threadIdx.x + k, threadIdx.x - k);
    const float2 xx = make_float2(x, x);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = __ffma2_rn(a[k], xx, a[(k + 1) & 15]);
    }
    float t = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += a[k].x + a[k].y;
    if (t == 1.2345f) out[0] = t;
}

template <typename F> float time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s, %d SMs, max clock %d MHz, smem optin %zu\n", prop.name, sms, clk_khz / 1000, prop.sharedMemPerBlockOptin);
    const size_t slab_bytes = (size_t)ROWS * 128;
    const size_t all_rows = (size_t)40 * 1029 * 8;  // rows of the S4 value tensor
    float *big;
    CK(cudaMalloc(&big, all_rows * 128));
    CK(cudaMemset(big, 0, all_rows * 128));
    float *out;
    CK(cudaMalloc(&out, 1024));
    CK(cudaFuncSetAttribute(lds_groups, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes));
    CK(cudaFuncSetAttribute(lds_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes));
    CK(cudaFuncSetAttribute(atoms_int, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes));
    CK(cudaFuncSetAttribute(atoms_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes));
    CK(cudaFuncSetAttribute(lds_sts_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes));
    CK(cudaFuncSetAttribute(ldg_groups, cudaFuncAttributePreferredSharedMemoryCarveout, 0));

    const double rows_per_launch = (double)sms * 1024 / 8 * ITERS;  // row accesses (128 B each) in the "groups" kernels
    auto report = [&](const char *name, float ms, double rows) {
        const double bytes = rows * 128.0;
        printf("%-44s %8.3f ms  %8.1f GB/s  %6.1f B/clk/SM@%dMHz  %7.2f Grow/s\n", name, ms, bytes / ms / 1e6,
               bytes / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, rows / ms / 1e6);
    };
    float ms;
    ms = time_ms([&] { lds_groups<<<sms, 1024, slab_bytes>>>(out, ITERS); });
    report("a1 LDS.128 gather, 8 lanes/row", ms, rows_per_launch);
    ms = time_ms([&] { lds_lanes<<<sms, 1024, slab_bytes>>>(out, ITERS); });
    report("a2 LDS.128 gather, lane/row rotated", ms, (double)sms * 1024 * (ITERS / 8));
    ms = time_ms([&] { ldg_groups<<<sms, 1024>>>((const float4 *)big, out, ITERS, ROWS, 1); });
    report("b  LDG.128 gather, per-CTA 131KB slab (L1)", ms, rows_per_launch);
    ms = time_ms([&] { ldg_groups<<<sms, 1024>>>((const float4 *)big, out, ITERS / 4, (int)all_rows, 0); });
    report("c  LDG.128 gather, 42MB random rows (L2)", ms, rows_per_launch / 4);
    ms = time_ms([&] { red_v4_groups<<<sms, 1024>>>(big, ITERS / 8, (int)all_rows); });
    report("d1 REDG.v4 scatter 42MB, 8 lanes/row", ms, rows_per_launch / 8);
    ms = time_ms([&] { red_v4_lanes<<<sms, 1024>>>(big, ITERS / 8, (int)all_rows); });
    report("d2 REDG.v4 scatter 42MB, lane/row", ms, (double)sms * 1024 * (ITERS / 8 / 8));
    ms = time_ms([&] { red_scalar_rows<<<sms, 1024>>>(big, ITERS / 8, (int)all_rows); });
    report("e  RED.f32 scatter 42MB, warp/row (reference)", ms, (double)sms * 32 * (ITERS / 8 / 4));
    ms = time_ms([&] { red_v4_groups<<<sms, 1024>>>(big, ITERS / 8, ROWS * 8); });
    report("d3 REDG.v4 scatter 1MB frame, 8 lanes/row", ms, rows_per_launch / 8);
    ms = time_ms([&] { atoms_int<<<sms, 1024, slab_bytes>>>((int *)out, ITERS); });
    report("f1 ATOMS.ADD.s32 smem rows, warp/row", ms, (double)sms * 32 * (ITERS / 4));
    ms = time_ms([&] { atoms_f32<<<sms, 1024, slab_bytes>>>(out, ITERS); });
    report("f2 atomicAdd f32 smem rows (CAS), warp/row", ms, (double)sms * 32 * (ITERS / 4));
    ms = time_ms([&] { lds_sts_rmw<<<sms, 1024, slab_bytes>>>(out, ITERS); });
    report("f3 LDS+STS.128 row RMW, 8 lanes/row", ms, rows_per_launch);
    ms = time_ms([&] { shfl_only<<<sms, 1024>>>(out, ITERS); });
    printf("%-44s %8.3f ms  %8.2f warp-SHFL/clk/SM\n", "g  SHFL.BFLY only", ms, (double)32 * ITERS * 4 / (ms * 1e-3) / (clk_khz * 1e3));
    ms = time_ms([&] { ffma_only<<<sms, 1024>>>(out, ITERS, 1.0001f); });
    printf("%-44s %8.3f ms  %8.2f warp-FFMA/clk/SM\n", "h  FFMA only", ms, (double)32 * ITERS * 16 / (ms * 1e-3) / (clk_khz * 1e3));
    ms = time_ms([&] { ffma2_only<<<sms, 1024>>>(out, ITERS, 1.0001f); });
    printf("%-44s %8.3f ms  %8.2f warp-FFMA2/clk/SM (x2 FMAs each)\n", "h2 FFMA2 only", ms, (double)32 * ITERS * 16 / (ms * 1e-3) / (clk_khz * 1e3));
    return 0;
}
