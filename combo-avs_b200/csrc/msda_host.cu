// msda_host.cu -- host-buffer entry point of the C ABI: forward + backward of the core on tensors that live in (pinned)
// HOST memory.  This is the end-to-end form of the reference-facing call (bench.py's `e2e` leg): frames are independent
// in this op (ms_deform_im2col_cuda.cuh:268-274: the batch index only selects the value slab), so the batch is cut into
// frame chunks that flow through three CUDA streams -- host->device copies, kernels, device->host copies -- with
// `slots` staging buffers in rotation.  PCIe then runs full duplex (measured on the B200 box: 55 GB/s H2D alone, 58 GB/s
// D2H alone, 46-50 GB/s each way at once; profiles/r01t_pcie_probe.txt) and the kernels hide under the copies.  The whole loop
// is native: at 20 chunks per step the per-call overhead of a Python loop would be the bottleneck.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/msda_b200.h"
#include "msda_kernels.h"

namespace {

constexpr int kMaxSlots = 4;

struct Slot {
    void *value = nullptr, *loc = nullptr, *attn = nullptr, *gout = nullptr;      // inputs of one chunk
    void *out = nullptr, *gvalue = nullptr, *gloc = nullptr, *gattn = nullptr;   // results of one chunk
    void *ws = nullptr;
    cudaEvent_t ready = nullptr, ready2 = nullptr, fwd_done = nullptr, bwd_done = nullptr, drained = nullptr;
    bool used = false;
};

int elem(int dtype) { return dtype == MSDA_F32 ? 4 : dtype == MSDA_BF16 ? 2 : dtype == MSDA_F64 ? 8 : 0; }

}  // namespace

struct msda_host_pipeline {
    int device = 0, chunk = 4, slots = 3;
    cudaStream_t s_in = nullptr, s_in2 = nullptr, s_run = nullptr, s_out = nullptr;
    Slot slot[kMaxSlots];
    size_t cap_in[4] = {0, 0, 0, 0}, cap_out[4] = {0, 0, 0, 0}, cap_ws = 0;  // bytes per slot currently allocated
    int64_t *geo_dev = nullptr;  // spatial_shapes followed by level_start_index
    int geo_levels = 0;
    std::vector<int64_t> geo_host;
    char error[256] = "";
};

namespace {

int fail(msda_host_pipeline *p, int code, const char *what, cudaError_t e)
{
    snprintf(p->error, sizeof(p->error), "%s: %s", what, e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
    return code;
}

// (re)allocate one device buffer per slot when the chunk grew
int ensure(msda_host_pipeline *p, void *Slot::*field, size_t &cap, size_t need)
{
    if (need <= cap) return MSDA_OK;
    for (int i = 0; i < p->slots; ++i) {
        if (p->slot[i].*field) cudaFree(p->slot[i].*field);
        p->slot[i].*field = nullptr;
        cudaError_t e = cudaMalloc(&(p->slot[i].*field), need);
        if (e != cudaSuccess) return fail(p, (int)e, "cudaMalloc(staging)", e);
    }
    cap = need;
    return MSDA_OK;
}

}  // namespace

extern "C" {

const char *msda_host_pipeline_error(const msda_host_pipeline *p) { return p ? p->error : "NULL pipeline"; }

int msda_host_pipeline_create(msda_host_pipeline **out, int device, int chunk_frames, int slots)
{
    if (!out) return MSDA_ERR_NULL;
    *out = nullptr;
    if (chunk_frames < 1 || slots < 2 || slots > kMaxSlots) return MSDA_ERR_DIM;
    msda_host_pipeline *p = new (std::nothrow) msda_host_pipeline();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->device = device;
    p->chunk = chunk_frames;
    p->slots = slots;
    int prev = -1;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess && prev != device) e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_in2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < slots && e == cudaSuccess; ++i) {
        Slot &s = p->slot[i];
        e = cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.ready2, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.fwd_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.bwd_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.drained, cudaEventDisableTiming);
    }
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        msda_host_pipeline_destroy(p);
        return (int)e;
    }
    *out = p;
    return MSDA_OK;
}

int msda_host_pipeline_destroy(msda_host_pipeline *p)
{
    if (!p) return MSDA_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != p->device) cudaSetDevice(p->device);
    if (p->s_out) cudaStreamSynchronize(p->s_out);
    if (p->s_run) cudaStreamSynchronize(p->s_run);
    if (p->s_in) cudaStreamSynchronize(p->s_in);
    if (p->s_in2) cudaStreamSynchronize(p->s_in2);
    for (int i = 0; i < kMaxSlots; ++i) {
        Slot &s = p->slot[i];
        void *bufs[] = {s.value, s.loc, s.attn, s.gout, s.out, s.gvalue, s.gloc, s.gattn, s.ws};
        for (void *b : bufs)
            if (b) cudaFree(b);
        cudaEvent_t evs[] = {s.ready, s.ready2, s.fwd_done, s.bwd_done, s.drained};
        for (cudaEvent_t ev : evs)
            if (ev) cudaEventDestroy(ev);
    }
    if (p->geo_dev) cudaFree(p->geo_dev);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_in2) cudaStreamDestroy(p->s_in2);
    if (p->s_run) cudaStreamDestroy(p->s_run);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    delete p;
    return MSDA_OK;
}

int msda_forward_backward_host(msda_host_pipeline *p, int dtype, int loc_dtype, const void *value, const int64_t *shapes_host,
                               const int64_t *level_start_host, const void *sampling_loc, const void *attn_weight,
                               const void *grad_output, void *output, void *grad_value, void *grad_sampling_loc,
                               void *grad_attn_weight, int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                               uint64_t *h2d_bytes, uint64_t *d2h_bytes)
{
    if (!p) return MSDA_ERR_NULL;
    p->error[0] = 0;
    const int vb = elem(dtype), mb = elem(loc_dtype);
    if (!vb || !mb) return fail(p, MSDA_ERR_DTYPE, "unknown dtype", cudaSuccess);
    if (N < 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || L > msda::kMaxLevels || Lq <= 0 || P <= 0) return fail(p, MSDA_ERR_DIM, "bad dimensions", cudaSuccess);
    if (!shapes_host || !level_start_host) return fail(p, MSDA_ERR_NULL, "host geometry is required", cudaSuccess);
    if (N > 0 && (!value || !sampling_loc || !attn_weight || !grad_output || !output || !grad_value || !grad_sampling_loc || !grad_attn_weight))
        return fail(p, MSDA_ERR_NULL, "NULL host tensor", cudaSuccess);
    if (im2col_step <= 0) return fail(p, MSDA_ERR_STEP, "im2col_step must be positive", cudaSuccess);
    {   // the reference's batch check (ms_deform_attn_cuda.cu:55-57) applies to the whole batch, not to our chunks
        const int step = N < im2col_step ? N : im2col_step;
        if (N > 0 && N % step != 0) return fail(p, MSDA_ERR_STEP, "batch must divide im2col_step", cudaSuccess);
    }
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != p->device) cudaSetDevice(p->device);
    struct Restore {
        int prev, dev;
        ~Restore() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
    } restore{prev, p->device};

    // geometry on the device (the generic kernels read it there); refreshed only when it changes
    if (p->geo_levels != L || p->geo_host.size() != (size_t)3 * L || memcmp(p->geo_host.data(), shapes_host, sizeof(int64_t) * 2 * L) != 0 ||
        memcmp(p->geo_host.data() + 2 * L, level_start_host, sizeof(int64_t) * L) != 0) {
        cudaStreamSynchronize(p->s_run);
        if (p->geo_dev) cudaFree(p->geo_dev);
        p->geo_dev = nullptr;
        cudaError_t e = cudaMalloc(&p->geo_dev, sizeof(int64_t) * 3 * L);
        if (e != cudaSuccess) return fail(p, (int)e, "cudaMalloc(geometry)", e);
        p->geo_host.assign(shapes_host, shapes_host + 2 * L);
        p->geo_host.insert(p->geo_host.end(), level_start_host, level_start_host + L);
        e = cudaMemcpy(p->geo_dev, p->geo_host.data(), sizeof(int64_t) * 3 * L, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail(p, (int)e, "cudaMemcpy(geometry)", e);
        p->geo_levels = L;
    }
    const int64_t *shapes_dev = p->geo_dev, *start_dev = p->geo_dev + 2 * L;

    const size_t fv = (size_t)S * M * D * vb, fl = (size_t)Lq * M * L * P * 2 * mb, fa = (size_t)Lq * M * L * P * mb,
                 fo = (size_t)Lq * M * D * vb;  // bytes per frame
    const int chunk = p->chunk < N ? p->chunk : (N > 0 ? N : 1);
    const size_t wsb = msda_backward_workspace_bytes(dtype, loc_dtype, chunk, S, M, D, L, Lq, P);
    {
        // results may still be draining from a previous call into the old buffers: quiesce before any reallocation
        const size_t need_in[4] = {fv * chunk, fl * chunk, fa * chunk, fo * chunk};
        const size_t need_out[4] = {fo * chunk, fv * chunk, fl * chunk, fa * chunk};
        bool grow = wsb > p->cap_ws;
        for (int k = 0; k < 4; ++k) grow |= need_in[k] > p->cap_in[k] || need_out[k] > p->cap_out[k];
        if (grow) {
            cudaStreamSynchronize(p->s_in);
            cudaStreamSynchronize(p->s_in2);
            cudaStreamSynchronize(p->s_run);
            cudaStreamSynchronize(p->s_out);
        }
        void *Slot::*in_f[4] = {&Slot::value, &Slot::loc, &Slot::attn, &Slot::gout};
        void *Slot::*out_f[4] = {&Slot::out, &Slot::gvalue, &Slot::gloc, &Slot::gattn};
        for (int k = 0; k < 4; ++k) {
            int rc = ensure(p, in_f[k], p->cap_in[k], need_in[k]);
            if (rc == MSDA_OK) rc = ensure(p, out_f[k], p->cap_out[k], need_out[k]);
            if (rc != MSDA_OK) return rc;
        }
        if (wsb > 0) {
            int rc = ensure(p, &Slot::ws, p->cap_ws, wsb);
            if (rc != MSDA_OK) return rc;
        }
    }

    // MSDA_HOST_TRACE=1: per-chunk timeline (ms since the first copy was queued) on stderr -- debugging aid only
    static const bool trace = [] { const char *e = getenv("MSDA_HOST_TRACE"); return e && atoi(e) == 1; }();
    std::vector<cudaEvent_t> tev;
    cudaEvent_t t0 = nullptr;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        tev.push_back(ev);
    };
    if (trace) {
        cudaEventCreate(&t0);
        cudaEventRecord(t0, p->s_in);
    }
    uint64_t up = 0, down = 0;
    const char *hv = (const char *)value, *hl = (const char *)sampling_loc, *ha = (const char *)attn_weight, *hg = (const char *)grad_output;
    char *ho = (char *)output, *hgv = (char *)grad_value, *hgl = (char *)grad_sampling_loc, *hga = (char *)grad_attn_weight;
    // every early return below leaves copies of earlier chunks in flight between the caller's host buffers and the device: drain
    // all pipeline streams before the caller gets control back (it may free those buffers as soon as it sees the error)
    struct Drain {
        msda_host_pipeline *p;
        bool armed = true;
        ~Drain()
        {
            if (!armed) return;
            cudaStreamSynchronize(p->s_in);
            cudaStreamSynchronize(p->s_in2);
            cudaStreamSynchronize(p->s_run);
            cudaStreamSynchronize(p->s_out);
        }
    } drain{p};
    int ci = 0;
    for (int lo = 0; lo < N; lo += chunk, ++ci) {
        const int n = (N - lo) < chunk ? (N - lo) : chunk;
        Slot &s = p->slot[ci % p->slots];
        cudaError_t e = cudaSuccess;
        // ---- host -> device: the slot's inputs are free once the backward that read them last has finished
        // (MSDA_HOST_SPLIT_IN=1 sends locations + weights down a second copy stream: more host reads in flight)
        static const bool split_in = [] { const char *e = getenv("MSDA_HOST_SPLIT_IN"); return e && atoi(e) == 1; }();
        cudaStream_t s_meta = split_in ? p->s_in2 : p->s_in;
        if (s.used) e = cudaStreamWaitEvent(p->s_in, s.bwd_done, 0);
        if (s.used && split_in && e == cudaSuccess) e = cudaStreamWaitEvent(p->s_in2, s.bwd_done, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.value, hv + lo * fv, n * fv, cudaMemcpyHostToDevice, p->s_in);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.loc, hl + lo * fl, n * fl, cudaMemcpyHostToDevice, s_meta);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.attn, ha + lo * fa, n * fa, cudaMemcpyHostToDevice, s_meta);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.gout, hg + lo * fo, n * fo, cudaMemcpyHostToDevice, p->s_in);
        if (e == cudaSuccess) e = cudaEventRecord(s.ready, p->s_in);
        if (e == cudaSuccess && split_in) e = cudaEventRecord(s.ready2, p->s_in2);
        if (e != cudaSuccess) return fail(p, (int)e, "host->device copies", e);
        mark(p->s_in);
        up += (uint64_t)n * (fv + fl + fa + fo);
        // ---- kernels: need the inputs, and the slot's result buffers drained by the previous round's device->host copies
        e = cudaStreamWaitEvent(p->s_run, s.ready, 0);
        if (e == cudaSuccess && split_in) e = cudaStreamWaitEvent(p->s_run, s.ready2, 0);
        if (e == cudaSuccess && s.used) e = cudaStreamWaitEvent(p->s_run, s.drained, 0);
        if (e != cudaSuccess) return fail(p, (int)e, "cudaStreamWaitEvent", e);
        int step = n < im2col_step ? n : im2col_step;
        while (n % step) --step;
        int rc = msda_forward(dtype, loc_dtype, s.value, shapes_dev, start_dev, s.loc, s.attn, s.out, n, S, M, D, L, Lq, P, step, shapes_host,
                              level_start_host, p->device, p->s_run);
        if (rc != MSDA_OK) {
            snprintf(p->error, sizeof(p->error), "msda_forward: %s", msda_last_error());
            return rc;
        }
        e = cudaEventRecord(s.fwd_done, p->s_run);
        if (e != cudaSuccess) return fail(p, (int)e, "cudaEventRecord", e);
        mark(p->s_run);
        rc = msda_backward(dtype, loc_dtype, s.value, shapes_dev, start_dev, s.loc, s.attn, s.gout, s.gvalue, s.gloc, s.gattn, n, S, M, D, L,
                           Lq, P, step, shapes_host, level_start_host, s.ws, p->cap_ws, p->device, p->s_run);
        if (rc != MSDA_OK) {
            snprintf(p->error, sizeof(p->error), "msda_backward: %s", msda_last_error());
            return rc;
        }
        e = cudaEventRecord(s.bwd_done, p->s_run);
        mark(p->s_run);
        // ---- device -> host: the forward result leaves while the backward still runs
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s_out, s.fwd_done, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ho + lo * fo, s.out, n * fo, cudaMemcpyDeviceToHost, p->s_out);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s_out, s.bwd_done, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hgv + lo * fv, s.gvalue, n * fv, cudaMemcpyDeviceToHost, p->s_out);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hgl + lo * fl, s.gloc, n * fl, cudaMemcpyDeviceToHost, p->s_out);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hga + lo * fa, s.gattn, n * fa, cudaMemcpyDeviceToHost, p->s_out);
        if (e == cudaSuccess) e = cudaEventRecord(s.drained, p->s_out);
        if (e != cudaSuccess) return fail(p, (int)e, "device->host copies", e);
        mark(p->s_out);
        down += (uint64_t)n * (fo + fv + fl + fa);
        s.used = true;
    }
    cudaError_t e = cudaStreamSynchronize(p->s_out);  // every result is on the host when the call returns
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_run);
    if (e != cudaSuccess) return fail(p, (int)e, "pipeline execution", e);
    drain.armed = false;
    if (trace) {
        for (size_t i = 0; i + 3 < tev.size() + 1 && i + 3 <= tev.size(); i += 4) {
            float a, b, c, d;
            cudaEventElapsedTime(&a, t0, tev[i]);
            cudaEventElapsedTime(&b, t0, tev[i + 1]);
            cudaEventElapsedTime(&c, t0, tev[i + 2]);
            cudaEventElapsedTime(&d, t0, tev[i + 3]);
            fprintf(stderr, "[msda host] chunk %2zu: h2d done %.3f  fwd done %.3f  bwd done %.3f  d2h done %.3f ms\n", i / 4, a, b, c, d);
        }
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
        cudaEventDestroy(t0);
    }
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    return MSDA_OK;
}

}  // extern "C"
