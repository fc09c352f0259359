// msda_api.cu -- the C ABI of libmsda_b200.so (include/msda_b200.h): argument validation, the checks the
// reference host code makes (ops/src/cuda/ms_deform_attn_cuda.cu:33-57, 98-121), variant selection and dispatch.
// No torch / ATen types cross this boundary; the caller owns every buffer.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/msda_b200.h"
#include "msda_kernels.h"

namespace msda {

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

namespace {

thread_local char t_error[512] = "";
std::atomic<const char *> g_last_variant[2] = {{"none"}, {"none"}};  // process-wide: backward runs on the autograd thread

enum FwdVariant { kFwdAuto = 0, kFwdGeneric, kFwdWarp32, kFwdSlab32, kFwdCount };
enum BwdVariant { kBwdAuto = 0, kBwdGeneric, kBwdWarp32, kBwdSorted32, kBwdCount };
const char *const kFwdNames[kFwdCount] = {"auto", "generic", "warp32", "slab32"};
const char *const kBwdNames[kBwdCount] = {"auto", "generic", "warp32", "sorted32"};
std::atomic<int> g_forced[2] = {{0}, {0}};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail((int)e, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

struct DeviceInfo {
    int sm_count = 0;
    int smem_optin = 0;
    bool ok = false;
};

// Device properties are immutable: cache them per ordinal (read-only after first use).
int device_info(int device, DeviceInfo &out)
{
    static DeviceInfo cache[64];
    static std::mutex mu;
    if (device < 0 || device >= 64) return fail(MSDA_ERR_DIM, "device ordinal %d out of range", device);
    std::lock_guard<std::mutex> lock(mu);
    if (!cache[device].ok) {
        cudaError_t e = cudaDeviceGetAttribute(&cache[device].sm_count, cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(sm count)");
        e = cudaDeviceGetAttribute(&cache[device].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(smem optin)");
        int major = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(cc major)");
        if (major != 10)
            return fail((int)cudaErrorNoKernelImageForDevice,
                        "libmsda_b200 holds sm_100a code only; device %d has compute capability %d.x", device, major);
        cache[device].ok = true;
    }
    out = cache[device];
    return MSDA_OK;
}

int elem_size(int dtype) { return dtype == MSDA_F32 ? 4 : dtype == MSDA_BF16 ? 2 : dtype == MSDA_F64 ? 8 : 0; }

int fill_problem(Problem &p, int dtype, int loc_dtype, const void *value, const int64_t *shapes, const int64_t *start,
                 const void *loc, const void *attn, int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                 const int64_t *shapes_host, const int64_t *start_host, int device)
{
    if (elem_size(dtype) == 0) return fail(MSDA_ERR_DTYPE, "unknown dtype %d", dtype);
    if (!((dtype == MSDA_BF16 && (loc_dtype == MSDA_BF16 || loc_dtype == MSDA_F32)) || (dtype != MSDA_BF16 && loc_dtype == dtype)))
        return fail(MSDA_ERR_DTYPE, "unsupported (dtype=%d, loc_dtype=%d): locations/weights must match value, or be f32 with bf16 values",
                    dtype, loc_dtype);
    if (N < 0 || S < 0 || Lq < 0 || M <= 0 || D <= 0 || L <= 0 || P <= 0)
        return fail(MSDA_ERR_DIM, "bad dimensions N=%d S=%d M=%d D=%d L=%d Lq=%d P=%d", N, S, M, D, L, Lq, P);
    const long long lim = 0x7fffffffLL;
    if ((long long)S * M * D > lim || (long long)Lq * M * L * P * 2 > lim || (long long)N * Lq * M > lim / 64)
        return fail(MSDA_ERR_DIM, "problem too large for 32-bit per-frame indexing");
    if (im2col_step <= 0) return fail(MSDA_ERR_STEP, "im2col_step must be positive, got %d", im2col_step);
    const int step = N < im2col_step ? N : im2col_step;
    if (N > 0 && N % step != 0) return fail(MSDA_ERR_STEP, "batch(%d) must divide im2col_step(%d)", N, step);
    const bool empty = (N == 0 || Lq == 0);
    if (!empty && (!value || !shapes || !start || !loc || !attn)) {
        if (!(S == 0 && !value && shapes && start && loc && attn)) return fail(MSDA_ERR_NULL, "NULL input pointer");
    }
    DeviceInfo di;
    int rc = device_info(device, di);
    if (rc != MSDA_OK) return rc;
    p.dtype = dtype;
    p.loc_dtype = loc_dtype;
    p.N = N; p.S = S; p.M = M; p.D = D; p.L = L; p.Lq = Lq; p.P = P;
    p.value = value; p.shapes = shapes; p.start = start; p.loc = loc; p.attn = attn;
    p.sm_count = di.sm_count;
    p.smem_optin = di.smem_optin;
    p.device = device;
    p.have_host = shapes_host != nullptr && start_host != nullptr && L <= kMaxLevels;
    if (p.have_host) {
        long long total = 0;
        for (int l = 0; l < L; ++l) {
            const int64_t H = shapes_host[2 * l], W = shapes_host[2 * l + 1];
            if (H <= 0 || W <= 0 || H > 1 << 20 || W > 1 << 20) return fail(MSDA_ERR_DIM, "level %d has shape %lld x %lld", l, (long long)H, (long long)W);
            if (start_host[l] < 0 || start_host[l] + H * W > S)
                return fail(MSDA_ERR_DIM, "level %d (start %lld, %lld x %lld) exceeds S=%d", l, (long long)start_host[l], (long long)H, (long long)W, S);
            p.table.H[l] = (int)H;
            p.table.W[l] = (int)W;
            p.table.start[l] = (int)start_host[l];
            total += H * W;
        }
        (void)total;
    }
    return MSDA_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int device)
    {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        if (prev != device) {
            e = cudaSetDevice(device);
            switched = e == cudaSuccess;
        }
        return e;
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

}  // namespace
}  // namespace msda

using namespace msda;

extern "C" {

int msda_version(void) { return MSDA_VERSION; }

const char *msda_last_error(void) { return t_error; }

uint64_t msda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char *msda_variant_names(int pass_)
{
    return pass_ == 0 ? "auto,generic,warp32,slab32" : "auto,generic,warp32,sorted32";
}

const char *msda_last_variant(int pass_) { return (pass_ == 0 || pass_ == 1) ? g_last_variant[pass_].load() : "none"; }

int msda_select_variant(int pass_, const char *name)
{
    if (pass_ != 0 && pass_ != 1) return fail(MSDA_ERR_VARIANT, "pass must be 0 (forward) or 1 (backward)");
    if (!name) return fail(MSDA_ERR_NULL, "variant name is NULL");
    const char *const *names = pass_ == 0 ? kFwdNames : kBwdNames;
    const int count = pass_ == 0 ? (int)kFwdCount : (int)kBwdCount;
    for (int i = 0; i < count; ++i)
        if (strcmp(name, names[i]) == 0) {
            g_forced[pass_].store(i);
            return MSDA_OK;
        }
    return fail(MSDA_ERR_VARIANT, "unknown %s variant '%s' (have: %s)", pass_ == 0 ? "forward" : "backward", name, msda_variant_names(pass_));
}

int msda_forward(int dtype, int loc_dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const void *sampling_loc, const void *attn_weight, void *output, int N, int S, int M, int D, int L, int Lq, int P,
                 int im2col_step, const int64_t *shapes_host, const int64_t *level_start_host, int device, void *stream)
{
    Problem p;
    int rc = fill_problem(p, dtype, loc_dtype, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, N, S, M, D, L, Lq,
                          P, im2col_step, shapes_host, level_start_host, device);
    if (rc != MSDA_OK) return rc;
    if (N == 0 || Lq == 0) return MSDA_OK;
    if (!output) return fail(MSDA_ERR_NULL, "output is NULL");
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;

    int v = g_forced[0].load();
    if (v == kFwdAuto) v = slab32_supported(p) ? kFwdSlab32 : warp32_supported(p) ? kFwdWarp32 : kFwdGeneric;
    if (S == 0) v = kFwdGeneric;
    switch (v) {
    case kFwdSlab32:
        if (!slab32_supported(p))
            return fail(MSDA_ERR_VARIANT, "forward variant slab32 needs f32, D=32, host geometry and a padded slab that fits shared memory");
        e = fwd_slab32(p, output, st);
        break;
    case kFwdWarp32:
        if (!warp32_supported(p)) return fail(MSDA_ERR_VARIANT, "forward variant warp32 needs f32, D=32, L<=%d", kMaxLevels);
        e = fwd_warp32(p, output, st);
        break;
    default:
        v = kFwdGeneric;
        e = fwd_generic(p, output, st);
        break;
    }
    g_last_variant[0].store(kFwdNames[v]);
    if (e != cudaSuccess) return cuda_fail(e, "msda_forward launch");
    return MSDA_OK;
}

size_t msda_backward_workspace_bytes(int dtype, int loc_dtype, int N, int S, int M, int D, int L, int Lq, int P)
{
    (void)loc_dtype; (void)L; (void)Lq; (void)P;
    if (N <= 0 || S <= 0 || M <= 0 || D <= 0) return 0;
    return bwd_generic_workspace(dtype, N, S, M, D);
}

int msda_backward(int dtype, int loc_dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                  const void *sampling_loc, const void *attn_weight, const void *grad_output, void *grad_value,
                  void *grad_sampling_loc, void *grad_attn_weight, int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                  const int64_t *shapes_host, const int64_t *level_start_host, void *workspace, size_t workspace_bytes, int device,
                  void *stream)
{
    Problem p;
    int rc = fill_problem(p, dtype, loc_dtype, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, N, S, M, D, L, Lq,
                          P, im2col_step, shapes_host, level_start_host, device);
    if (rc != MSDA_OK) return rc;
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_val = (size_t)N * S * M * D;
    if (N == 0 || Lq == 0) {  // no samples: grad_value is all zero, the other two are empty
        if (n_val > 0) {
            if (!grad_value) return fail(MSDA_ERR_NULL, "grad_value is NULL");
            e = cudaMemsetAsync(grad_value, 0, n_val * elem_size(dtype), st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_value)");
        }
        return MSDA_OK;
    }
    if (!grad_output || !grad_sampling_loc || !grad_attn_weight || (n_val > 0 && !grad_value))
        return fail(MSDA_ERR_NULL, "NULL gradient pointer");
    const size_t need = msda_backward_workspace_bytes(dtype, loc_dtype, N, S, M, D, L, Lq, P);
    if (need > 0 && (!workspace || workspace_bytes < need))
        return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);

    BackwardIO io{grad_output, grad_value, grad_sampling_loc, grad_attn_weight, workspace, workspace_bytes};
    int v = g_forced[1].load();
    if (v == kBwdAuto) v = sorted32_supported(p) ? kBwdSorted32 : warp32_supported(p) ? kBwdWarp32 : kBwdGeneric;
    if (S == 0) v = kBwdGeneric;
    switch (v) {
    case kBwdSorted32:
        if (!sorted32_supported(p))
            return fail(MSDA_ERR_VARIANT, "backward variant sorted32 needs f32, D=32, host geometry, Lq*P <= 4608, (H+1)*(W+1) <= 1024 per level");
        e = bwd_sorted32(p, io, st);
        break;
    case kBwdWarp32:
        if (!warp32_supported(p)) return fail(MSDA_ERR_VARIANT, "backward variant warp32 needs f32, D=32, L<=%d", kMaxLevels);
        e = bwd_warp32(p, io, st);
        break;
    default:
        v = kBwdGeneric;
        e = bwd_generic(p, io, st);
        break;
    }
    g_last_variant[1].store(kBwdNames[v]);
    if (e != cudaSuccess) return cuda_fail(e, "msda_backward launch");
    return MSDA_OK;
}


/* ---- fused module prologue (ops/modules/ms_deform_attn.py:102-116) ------------------------------------------ */

int msda_fused_supported(int dtype, int N, int S, int M, int D, int L, int Lq, int P, const int64_t *shapes_host,
                         const int64_t *level_start_host, int device)
{
    if (dtype != MSDA_F32 || !shapes_host || !level_start_host || N <= 0 || Lq <= 0 || S <= 0) return 0;
    Problem p;
    int dummy = 0;
    // pointers are only inspected for NULL-ness by the support queries
    int rc = fill_problem(p, dtype, dtype, &dummy, (const int64_t *)&dummy, (const int64_t *)&dummy, &dummy, &dummy, N, S, M, D, L, Lq, P, 128,
                          shapes_host, level_start_host, device);
    if (rc != MSDA_OK) return 0;
    p.ref = &dummy;
    p.attn_out = &dummy;
    return slab32_supported(p) && sorted32_supported(p) ? 1 : 0;
}

int msda_fused_forward(int dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *reference_points, const void *sampling_offsets, const void *attention_logits,
                       const unsigned char *padding_mask, void *output, void *attention_weights_out, int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                       const int64_t *shapes_host, const int64_t *level_start_host, int device, void *stream)
{
    Problem p;
    int rc = fill_problem(p, dtype, dtype, value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, N, S, M, D, L, Lq,
                          P, im2col_step, shapes_host, level_start_host, device);
    if (rc != MSDA_OK) return rc;
    if (N == 0 || Lq == 0) return MSDA_OK;
    if (!output || !reference_points || !attention_weights_out) return fail(MSDA_ERR_NULL, "NULL pointer in fused forward");
    p.ref = reference_points;
    p.attn_out = attention_weights_out;
    p.mask = padding_mask;
    if (dtype != MSDA_F32 || !slab32_supported(p))
        return fail(MSDA_ERR_VARIANT, "fused forward needs f32, D=32, L*P %% 4 == 0, host geometry and a padded slab that fits shared memory");
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    e = fwd_slab32(p, output, (cudaStream_t)stream);
    g_last_variant[0].store("slab32+prologue");
    if (e != cudaSuccess) return cuda_fail(e, "msda_fused_forward launch");
    return MSDA_OK;
}

int msda_fused_backward(int dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *reference_points, const void *sampling_offsets, const void *attention_weights,
                        const unsigned char *padding_mask, const void *grad_output, void *grad_value, void *grad_sampling_offsets, void *grad_attention_logits, int N,
                        int S, int M, int D, int L, int Lq, int P, int im2col_step, const int64_t *shapes_host,
                        const int64_t *level_start_host, int device, void *stream)
{
    Problem p;
    int rc = fill_problem(p, dtype, dtype, value, spatial_shapes, level_start_index, sampling_offsets, attention_weights, N, S, M, D, L, Lq,
                          P, im2col_step, shapes_host, level_start_host, device);
    if (rc != MSDA_OK) return rc;
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_val = (size_t)N * S * M * D;
    if (N == 0 || Lq == 0) {
        if (n_val > 0 && grad_value) {
            e = cudaMemsetAsync(grad_value, 0, n_val * elem_size(dtype), st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_value)");
        }
        return MSDA_OK;
    }
    if (!grad_output || !grad_sampling_offsets || !grad_attention_logits || !reference_points || (n_val > 0 && !grad_value))
        return fail(MSDA_ERR_NULL, "NULL pointer in fused backward");
    p.ref = reference_points;
    p.mask = padding_mask;
    if (dtype != MSDA_F32 || !sorted32_supported(p))
        return fail(MSDA_ERR_VARIANT, "fused backward needs f32, D=32, host geometry, Lq*P <= 4608, (H+1)*(W+1) <= 1024 per level");
    // the sorted backward leaves d(out)/d(softmaxed weight) in grad_attention_logits; the softmax backward then runs in place
    BackwardIO io{grad_output, grad_value, grad_sampling_offsets, grad_attention_logits, nullptr, 0};
    e = bwd_sorted32(p, io, st);
    if (e == cudaSuccess)
        e = softmax_backward_rows((const float *)attention_weights, (float *)grad_attention_logits, (long long)N * Lq * M, L * P, p.sm_count, st);
    if (e == cudaSuccess && padding_mask) e = zero_masked_rows((float *)grad_value, padding_mask, (long long)N * S, M * D, p.sm_count, st);
    g_last_variant[1].store("sorted32+prologue");
    if (e != cudaSuccess) return cuda_fail(e, "msda_fused_backward launch");
    return MSDA_OK;
}

/* ---- encoder-layer Linear on the tensor cores (msdeformattn.py:119-134, ms_deform_attn.py:105-128) ------------------ */

int msda_linear_supported(int M, int N, int K) { return linear_supported(M, N, K) ? 1 : 0; }

int msda_linear_forward(const void *x, const void *weight, const void *bias, void *y, int M, int N, int K, int flags, int device, void *stream)
{
    if (M < 0 || N <= 0 || K <= 0) return fail(MSDA_ERR_DIM, "bad dimensions M=%d N=%d K=%d", M, N, K);
    if (M == 0) return MSDA_OK;
    if (!x || !weight || !y) return fail(MSDA_ERR_NULL, "NULL pointer in msda_linear_forward");
    DeviceInfo di;
    int rc = device_info(device, di);
    if (rc != MSDA_OK) return rc;
    if (!linear_supported(M, N, K))
        return fail(MSDA_ERR_VARIANT, "msda_linear_forward needs K %% 4 == 0, N %% 16 == 0 and a driver with cuTensorMapEncodeTiled (M=%d N=%d K=%d)", M, N, K);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15)
        return fail(MSDA_ERR_DIM, "msda_linear_forward needs 16-byte aligned pointers");
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    e = linear_tf32((const float *)x, (const float *)weight, (const float *)bias, (float *)y, M, N, K, flags & MSDA_LINEAR_RELU,
                    (flags & MSDA_LINEAR_TF32_ONLY) ? 0 : 1, 0, device, di.sm_count, di.smem_optin, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "msda_linear_forward launch");
    return MSDA_OK;
}

int msda_linear_wgrad_supported(int M, int N, int K) { return linear_wgrad_supported(M, N, K) ? 1 : 0; }

int msda_linear_wgrad(const void *grad_y, const void *x, void *grad_weight, int M, int N, int K, int device, void *stream)
{
    if (M < 0 || N <= 0 || K <= 0) return fail(MSDA_ERR_DIM, "bad dimensions M=%d N=%d K=%d", M, N, K);
    if (!grad_weight) return fail(MSDA_ERR_NULL, "grad_weight is NULL");
    DeviceInfo di;
    int rc = device_info(device, di);
    if (rc != MSDA_OK) return rc;
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    if (M == 0) {  // no rows: the gradient is zero
        e = cudaMemsetAsync(grad_weight, 0, (size_t)N * K * sizeof(float), (cudaStream_t)stream);
        return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "cudaMemsetAsync(grad_weight)");
    }
    if (!grad_y || !x) return fail(MSDA_ERR_NULL, "NULL pointer in msda_linear_wgrad");
    if (!linear_wgrad_supported(M, N, K))
        return fail(MSDA_ERR_VARIANT, "msda_linear_wgrad needs N %% 4 == 0, K %% 32 == 0 and a driver with cuTensorMapEncodeTiled (M=%d N=%d K=%d)", M, N, K);
    if ((reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_weight)) & 15)
        return fail(MSDA_ERR_DIM, "msda_linear_wgrad needs 16-byte aligned pointers");
    e = linear_tf32((const float *)grad_y, (const float *)x, nullptr, (float *)grad_weight, N, K, M, 0, 1, 1, device, di.sm_count, di.smem_optin,
                    (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "msda_linear_wgrad launch");
    return MSDA_OK;
}

int msda_layernorm_supported(int C) { return layernorm_supported(C) ? 1 : 0; }

size_t msda_layernorm_backward_workspace_bytes(int M, int C, int device)
{
    DeviceInfo di;
    if (M <= 0 || !layernorm_supported(C) || device_info(device, di) != MSDA_OK) return 0;
    return layernorm_backward_workspace(M, C, di.sm_count);
}

int msda_add_layernorm_forward(const void *a, const void *b, const void *gamma, const void *beta, void *y, void *mean, void *rstd,
                               int M, int C, float eps, int device, void *stream)
{
    if (M < 0 || !layernorm_supported(C)) return fail(MSDA_ERR_DIM, "msda_add_layernorm_forward needs M >= 0 and C in {128, 256, 384, 512} (M=%d C=%d)", M, C);
    if (M == 0) return MSDA_OK;
    if (!a || !gamma || !beta || !y || !mean || !rstd) return fail(MSDA_ERR_NULL, "NULL pointer in msda_add_layernorm_forward");
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
         reinterpret_cast<uintptr_t>(y)) & 15)
        return fail(MSDA_ERR_DIM, "msda_add_layernorm_forward needs 16-byte aligned pointers");
    DeviceInfo di;
    int rc = device_info(device, di);
    if (rc != MSDA_OK) return rc;
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    e = add_layernorm_forward((const float *)a, (const float *)b, (const float *)gamma, (const float *)beta, (float *)y, (float *)mean, (float *)rstd, M, C,
                              eps, di.sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "msda_add_layernorm_forward launch");
    return MSDA_OK;
}

int msda_layernorm_backward(const void *grad_y, const void *a, const void *b, const void *gamma, const void *mean, const void *rstd,
                            void *grad_x, void *grad_gamma, void *grad_beta, void *workspace, size_t workspace_bytes,
                            int M, int C, int device, void *stream)
{
    if (M < 0 || !layernorm_supported(C)) return fail(MSDA_ERR_DIM, "msda_layernorm_backward needs M >= 0 and C in {128, 256, 384, 512} (M=%d C=%d)", M, C);
    if (!grad_gamma || !grad_beta) return fail(MSDA_ERR_NULL, "NULL gradient pointer in msda_layernorm_backward");
    DeviceInfo di;
    int rc = device_info(device, di);
    if (rc != MSDA_OK) return rc;
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    if (M == 0) {  // no rows: the parameter gradients are zero
        e = cudaMemsetAsync(grad_gamma, 0, (size_t)C * sizeof(float), (cudaStream_t)stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(grad_beta, 0, (size_t)C * sizeof(float), (cudaStream_t)stream);
        return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "cudaMemsetAsync(grad_gamma / grad_beta)");
    }
    if (!grad_y || !a || !gamma || !mean || !rstd || !grad_x) return fail(MSDA_ERR_NULL, "NULL pointer in msda_layernorm_backward");
    if ((reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(gamma) |
         reinterpret_cast<uintptr_t>(grad_x) | reinterpret_cast<uintptr_t>(workspace)) & 15)
        return fail(MSDA_ERR_DIM, "msda_layernorm_backward needs 16-byte aligned pointers");
    const size_t need = layernorm_backward_workspace(M, C, di.sm_count);
    if (!workspace || workspace_bytes < need)
        return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    e = layernorm_backward((const float *)grad_y, (const float *)a, (const float *)b, (const float *)gamma, (const float *)mean, (const float *)rstd,
                           (float *)grad_x, (float *)grad_gamma, (float *)grad_beta, (float *)workspace, M, C, di.sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "msda_layernorm_backward launch");
    return MSDA_OK;
}

int msda_colsum_supported(int N) { return colsum_supported(N) ? 1 : 0; }

size_t msda_colsum_workspace_bytes(int M, int N, int device)
{
    DeviceInfo di;
    if (M <= 0 || !colsum_supported(N) || device_info(device, di) != MSDA_OK) return 0;
    return colsum_workspace(M, N, di.sm_count);
}

int msda_colsum(const void *g, void *out, void *workspace, size_t workspace_bytes, int M, int N, int device, void *stream)
{
    if (M < 0 || !colsum_supported(N)) return fail(MSDA_ERR_DIM, "msda_colsum needs M >= 0, N %% 4 == 0 and N <= 4096 (M=%d N=%d)", M, N);
    if (!out) return fail(MSDA_ERR_NULL, "out is NULL");
    DeviceInfo di;
    int rc = device_info(device, di);
    if (rc != MSDA_OK) return rc;
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    if (M == 0) {
        e = cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), (cudaStream_t)stream);
        return e == cudaSuccess ? MSDA_OK : cuda_fail(e, "cudaMemsetAsync(out)");
    }
    if (!g) return fail(MSDA_ERR_NULL, "g is NULL");
    if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(workspace)) & 15) return fail(MSDA_ERR_DIM, "msda_colsum needs 16-byte aligned pointers");
    const size_t need = colsum_workspace(M, N, di.sm_count);
    if (!workspace || workspace_bytes < need)
        return fail(MSDA_ERR_WORKSPACE, "workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    e = colsum((const float *)g, (float *)out, (float *)workspace, M, N, di.sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "msda_colsum launch");
    return MSDA_OK;
}

}  // extern "C"
