/*
 * msda_oracle.c -- CPU restatement of COMBO-AVS's multi-scale deformable attention core.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (combo-avs_b200/, include/) may link,
 * import or call this file.  It is used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * tests/golden/ (npz files), which oracle/gen_golden.py produced by importing the reference's own
 * ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:53-72) in this container.
 *
 * What it restates (paths relative to /root/reference/models/modeling/pixel_decoder/ops/):
 *   forward  : src/cuda/ms_deform_im2col_cuda.cuh:242-304 (sample loop + in-range test)
 *              src/cuda/ms_deform_im2col_cuda.cuh:38-89   (bilinear fetch, zero outside the map)
 *   backward : src/cuda/ms_deform_im2col_cuda.cuh:306-408 (sample loop, per-point reductions)
 *              src/cuda/ms_deform_im2col_cuda.cuh:92-164  (corner scatter + location/weight partials)
 *   layout   : src/cuda/ms_deform_attn_cuda.cu:45-53      (N,S,M,D / L / Lq / P read-out)
 *
 * The restatement is organised per *sample* (n,q,m,l,p) with an inner channel loop rather than per
 * output element as the reference kernels are; the arithmetic per element is the same.
 *
 * Two instantiations: _f64 (double throughout) and _f32 (float throughout, pixel coordinate formed
 * with one fused multiply-add as the reference's compiled fp32 kernel does).
 * OpenMP over (n, m): every (n, m) pair owns a disjoint slice of every output, so no atomics.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int msda_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void msda_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* One bilinear tap set: the four neighbours of (h, w) with their weights; a neighbour outside the
 * map keeps weight but is flagged absent (reads as zero, receives no gradient). */
#define DEFINE_ORACLE(REAL, SUFFIX, PIXCOORD)                                                          \
    typedef struct {                                                                                   \
        int ok;          /* sample inside (-1, H) x (-1, W) */                                         \
        int64_t row[4];  /* flattened pixel index inside the level, -1 if absent */                    \
        REAL wt[4];      /* hh*hw, hh*lw, lh*hw, lh*lw */                                              \
        REAL lh, lw, hh, hw;                                                                           \
    } taps_##SUFFIX;                                                                                   \
                                                                                                       \
    static taps_##SUFFIX make_taps_##SUFFIX(REAL loc_x, REAL loc_y, int H, int W)                      \
    {                                                                                                  \
        taps_##SUFFIX t;                                                                               \
        const REAL h = PIXCOORD(loc_y, (REAL)H);                                                       \
        const REAL w = PIXCOORD(loc_x, (REAL)W);                                                       \
        t.ok = (h > (REAL)-1) && (w > (REAL)-1) && (h < (REAL)H) && (w < (REAL)W);                     \
        if (!t.ok) return t;                                                                           \
        const int y0 = (int)floor((double)h), x0 = (int)floor((double)w);                              \
        const int y1 = y0 + 1, x1 = x0 + 1;                                                            \
        t.lh = h - (REAL)y0;                                                                           \
        t.lw = w - (REAL)x0;                                                                           \
        t.hh = (REAL)1 - t.lh;                                                                         \
        t.hw = (REAL)1 - t.lw;                                                                         \
        t.wt[0] = t.hh * t.hw;                                                                         \
        t.wt[1] = t.hh * t.lw;                                                                         \
        t.wt[2] = t.lh * t.hw;                                                                         \
        t.wt[3] = t.lh * t.lw;                                                                         \
        const int top = y0 >= 0, bot = y1 <= H - 1, lef = x0 >= 0, rig = x1 <= W - 1;                  \
        t.row[0] = (top && lef) ? (int64_t)y0 * W + x0 : -1;                                           \
        t.row[1] = (top && rig) ? (int64_t)y0 * W + x1 : -1;                                           \
        t.row[2] = (bot && lef) ? (int64_t)y1 * W + x0 : -1;                                           \
        t.row[3] = (bot && rig) ? (int64_t)y1 * W + x1 : -1;                                           \
        return t;                                                                                      \
    }                                                                                                  \
                                                                                                       \
    int msda_oracle_forward_##SUFFIX(const REAL *value, const int64_t *shapes, const int64_t *start,   \
                                     const REAL *loc, const REAL *attn, REAL *out, int N, int S,       \
                                     int M, int D, int L, int Lq, int P)                               \
    {                                                                                                  \
        if (N < 0 || S < 0 || M <= 0 || D <= 0 || L <= 0 || Lq < 0 || P <= 0) return -1;               \
        const int64_t pix_stride = (int64_t)M * D;                                                     \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                       \
        for (int n = 0; n < N; ++n)                                                                    \
            for (int m = 0; m < M; ++m) {                                                              \
                const REAL *vbase = value + (int64_t)n * S * pix_stride + (int64_t)m * D;              \
                for (int q = 0; q < Lq; ++q) {                                                         \
                    const int64_t qm = ((int64_t)n * Lq + q) * M + m;                                  \
                    REAL *o = out + qm * D;                                                            \
                    for (int c = 0; c < D; ++c) o[c] = 0;                                              \
                    for (int l = 0; l < L; ++l) {                                                      \
                        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                  \
                        const REAL *vl = vbase + start[l] * pix_stride;                                \
                        for (int p = 0; p < P; ++p) {                                                  \
                            const int64_t s = (qm * L + l) * P + p;                                    \
                            const taps_##SUFFIX t = make_taps_##SUFFIX(loc[2 * s], loc[2 * s + 1], H, W); \
                            if (!t.ok) continue;                                                       \
                            const REAL a = attn[s];                                                    \
                            for (int c = 0; c < D; ++c) {                                              \
                                REAL v[4];                                                             \
                                for (int k = 0; k < 4; ++k)                                            \
                                    v[k] = t.row[k] >= 0 ? vl[t.row[k] * pix_stride + c] : (REAL)0;    \
                                o[c] += (t.wt[0] * v[0] + t.wt[1] * v[1] + t.wt[2] * v[2] +            \
                                         t.wt[3] * v[3]) * a;                                          \
                            }                                                                          \
                        }                                                                              \
                    }                                                                                  \
                }                                                                                      \
            }                                                                                          \
        return 0;                                                                                      \
    }                                                                                                  \
                                                                                                       \
    int msda_oracle_backward_##SUFFIX(const REAL *value, const int64_t *shapes, const int64_t *start,  \
                                      const REAL *loc, const REAL *attn, const REAL *grad_out,         \
                                      REAL *grad_value, REAL *grad_loc, REAL *grad_attn, int N,        \
                                      int S, int M, int D, int L, int Lq, int P)                       \
    {                                                                                                  \
        if (N < 0 || S < 0 || M <= 0 || D <= 0 || L <= 0 || Lq < 0 || P <= 0) return -1;               \
        const int64_t pix_stride = (int64_t)M * D;                                                     \
        memset(grad_value, 0, sizeof(REAL) * (size_t)N * S * pix_stride);                              \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                       \
        for (int n = 0; n < N; ++n)                                                                    \
            for (int m = 0; m < M; ++m) {                                                              \
                const int64_t slab = (int64_t)n * S * pix_stride + (int64_t)m * D;                     \
                for (int q = 0; q < Lq; ++q) {                                                         \
                    const int64_t qm = ((int64_t)n * Lq + q) * M + m;                                  \
                    const REAL *g = grad_out + qm * D;                                                 \
                    for (int l = 0; l < L; ++l) {                                                      \
                        const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                  \
                        const REAL *vl = value + slab + start[l] * pix_stride;                         \
                        REAL *gvl = grad_value + slab + start[l] * pix_stride;                         \
                        for (int p = 0; p < P; ++p) {                                                  \
                            const int64_t s = (qm * L + l) * P + p;                                    \
                            REAL acc_x = 0, acc_y = 0, acc_a = 0;                                      \
                            const taps_##SUFFIX t = make_taps_##SUFFIX(loc[2 * s], loc[2 * s + 1], H, W); \
                            if (t.ok) {                                                                \
                                const REAL a = attn[s];                                                \
                                for (int c = 0; c < D; ++c) {                                          \
                                    const REAL ga = g[c] * a;                                          \
                                    REAL v[4];                                                         \
                                    for (int k = 0; k < 4; ++k) {                                      \
                                        if (t.row[k] >= 0) {                                           \
                                            v[k] = vl[t.row[k] * pix_stride + c];                      \
                                            gvl[t.row[k] * pix_stride + c] += t.wt[k] * ga;            \
                                        } else {                                                       \
                                            v[k] = 0;                                                  \
                                        }                                                              \
                                    }                                                                  \
                                    /* d(bilinear)/dh and /dw, absent neighbours contribute 0 */      \
                                    const REAL dh = -t.hw * v[0] - t.lw * v[1] + t.hw * v[2] + t.lw * v[3]; \
                                    const REAL dw = -t.hh * v[0] + t.hh * v[1] - t.lh * v[2] + t.lh * v[3]; \
                                    acc_x += (REAL)W * dw * ga;                                        \
                                    acc_y += (REAL)H * dh * ga;                                        \
                                    acc_a += g[c] * (t.wt[0] * v[0] + t.wt[1] * v[1] +                 \
                                                     t.wt[2] * v[2] + t.wt[3] * v[3]);                 \
                                }                                                                      \
                            }                                                                          \
                            grad_loc[2 * s] = acc_x;                                                   \
                            grad_loc[2 * s + 1] = acc_y;                                               \
                            grad_attn[s] = acc_a;                                                      \
                        }                                                                              \
                    }                                                                                  \
                }                                                                                      \
            }                                                                                          \
        return 0;                                                                                      \
    }

#define PIXCOORD_F64(x, n) ((x) * (n) - 0.5)
#define PIXCOORD_F32(x, n) fmaf((x), (n), -0.5f)

DEFINE_ORACLE(double, f64, PIXCOORD_F64)
DEFINE_ORACLE(float, f32, PIXCOORD_F32)
