/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product
 * path (fpqvar_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * Whole-function C port (pthreads over groups) of the reference's fake-quant functions, used
 * as the TIMED CPU baseline ("kind": "port").  It restates the same arithmetic as
 * oracle/oracle.py step by step; tests/test_oracle_port.py checks it bit for bit against the
 * numpy oracle, which in turn is pinned to fixtures produced by the reference's own Python
 * (tests/golden).  All file:line citations are relative to /root/reference/; "qu.py" =
 * models_fp_quant_transform_rotate/quant_utils.py.
 *
 *   port_fake_quant_*        fp_quant_e{1,2,3}_per_group[_cuda] qu.py:250-378,
 *                            fp6_quant_*_per_{group,token}_cuda qu.py:503-574
 *   port_signsplit_*         fp_quant_e1m2_neg_e2m1_pos_per_group[_cuda] qu.py:381-452,
 *                            fp6_quant_int_neg_e2m3_pos_* qu.py:577-646 (no global clip here;
 *                            it is the identity for finite data, qu.py:421-422)
 *   port_transform_rotate_quant  basic_var.py:263,266 `.mul(s)` + dense matmul with the
 *                            block-diagonal Q (only the non-zero 128x128 blocks are visited,
 *                            which favours the CPU), result rounded to fp16 as the autocast
 *                            GEMM output is, then fp_quant_e2_per_group_cuda qu.py:313-330.
 *
 * Rounding rules are the literal loops of scan_quant.c (quant/quant_kernel.cu:25-37 and
 * quantize_to_nearest_grid qu.py:224-230).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

typedef _Float16 f16;

static inline float scan_rule(float xv, const float *grid, int k)
{
    float best = 102400.0f, zv = 0.0f;
    for (int i = 0; i < k; ++i) {
        float d = fabsf(xv - grid[i]);
        if (d <= best) { best = d; zv = grid[i]; }
    }
    return zv;
}

static inline float argmin_rule(float xv, const float *grid, int k)
{
    float best = fabsf(xv - grid[0]);
    int arg = 0;
    for (int i = 1; i < k; ++i) {
        float d = fabsf(xv - grid[i]);
        if (!isnan(best) && (isnan(d) || d < best)) { best = d; arg = i; }
    }
    return grid[arg];
}

static inline float round_rule(float v, const float *grid, int k, int tie)
{
    return tie == 0 ? scan_rule(v, grid, k) : argmin_rule(v, grid, k);
}

/* torch.max semantics: NaN propagates */
static inline float max_nan(float a, float b) { return (isnan(a) || isnan(b)) ? NAN : (a > b ? a : b); }
static inline float clamp3(float x) { return x < -3.0f ? -3.0f : (x > 3.0f ? 3.0f : x); }

/* ---- a minimal static-schedule parallel-for over rows (this image has no libgomp) ------ */
int port_num_threads(void)
{
    const char *e = getenv("ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

typedef void (*range_fn)(const void *ctx, size_t begin, size_t end);
typedef struct { range_fn fn; const void *ctx; size_t begin, end; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->ctx, j->begin, j->end); return NULL; }

static void parallel_rows(range_fn fn, const void *ctx, size_t n)
{
    int t = port_num_threads();
    if ((size_t)t > n) t = n ? (int)n : 1;
    if (t <= 1) { fn(ctx, 0, n); return; }
    pthread_t th[256];
    job_t jobs[256];
    const size_t per = (n + (size_t)t - 1) / (size_t)t;
    int started = 0;
    for (int i = 0; i < t; ++i) {
        size_t b = per * (size_t)i, e = b + per > n ? n : b + per;
        if (b >= e) break;
        jobs[i] = (job_t){fn, ctx, b, e};
        if (pthread_create(&th[i], NULL, job_main, &jobs[i]) != 0) { fn(ctx, b, e); th[i] = 0; }
        started = i + 1;
    }
    for (int i = 0; i < started; ++i) if (th[i]) pthread_join(th[i], NULL);
}

/* ---- symmetric ---------------------------------------------------------------------- */
#define DEFINE_FAKE_QUANT(NAME, IN_T, OUT_T, RND_IN)                                              \
typedef struct { const IN_T *x; OUT_T *out; size_t row_len; const float *grid; int k; float gmax;  \
                 int tie, do_clamp3; } NAME##_ctx;                                                \
static void NAME##_range(const void *vc, size_t r0, size_t r1)                                    \
{                                                                                                 \
    const NAME##_ctx *c = (const NAME##_ctx *)vc;                                                 \
    const IN_T *x = c->x; OUT_T *out = c->out; const size_t row_len = c->row_len;                 \
    const float *grid = c->grid; const int k = c->k, tie = c->tie, do_clamp3 = c->do_clamp3;      \
    const float gmax = c->gmax;                                                                   \
    for (size_t r = r0; r < r1; ++r) {                                                            \
        const IN_T *xr = x + (size_t)r * row_len;                                                 \
        OUT_T *orow = out + (size_t)r * row_len;                                                  \
        float a = 0.0f;                                                                           \
        for (size_t i = 0; i < row_len; ++i) {                                                    \
            float f = (float)xr[i];                                                               \
            if (do_clamp3) f = clamp3(f);                                                         \
            a = max_nan(a, fabsf(f));                                                             \
        }                                                                                         \
        volatile float s = RND_IN(a / gmax);                    /* qu.py:320 */                   \
        for (size_t i = 0; i < row_len; ++i) {                                                    \
            float f = (float)xr[i];                                                               \
            if (do_clamp3) f = clamp3(f);                                                         \
            volatile float v = RND_IN(f / s);                   /* qu.py:321 */                   \
            float q = round_rule(v, grid, k, tie);              /* qu.py:323-326 */               \
            volatile float o = q * s;                           /* qu.py:328 */                   \
            orow[i] = (OUT_T)o;                                                                   \
        }                                                                                         \
    }                                                                                             \
}                                                                                                 \
void NAME(const IN_T *x, OUT_T *out, size_t n_rows, size_t row_len, const float *grid, int k,     \
          float gmax, int tie, int do_clamp3)                                                     \
{                                                                                                 \
    NAME##_ctx c = {x, out, row_len, grid, k, gmax, tie, do_clamp3};                              \
    parallel_rows(NAME##_range, &c, n_rows);                                                      \
}

#define RND_F32(e) ((float)(e))
#define RND_F16(e) ((float)(f16)(e))

DEFINE_FAKE_QUANT(port_fake_quant_f32_f32, float, float, RND_F32)
DEFINE_FAKE_QUANT(port_fake_quant_f32_f16, float, f16, RND_F32)
DEFINE_FAKE_QUANT(port_fake_quant_f16_f16, f16, f16, RND_F16)
DEFINE_FAKE_QUANT(port_fake_quant_f16_f32, f16, float, RND_F16)

/* ---- sign-split --------------------------------------------------------------------- */
#define DEFINE_SIGNSPLIT(NAME, IN_T, OUT_T, RND_IN)                                               \
typedef struct { const IN_T *x; OUT_T *out; size_t row_len; const float *gneg; int kn; float nmax; \
                 const float *gpos; int kp; float pmax; int tie; } NAME##_ctx;                    \
static void NAME##_range(const void *vc, size_t r0, size_t r1)                                    \
{                                                                                                 \
    const NAME##_ctx *c = (const NAME##_ctx *)vc;                                                 \
    const IN_T *x = c->x; OUT_T *out = c->out; const size_t row_len = c->row_len;                 \
    const float *gneg = c->gneg, *gpos = c->gpos; const int kn = c->kn, kp = c->kp, tie = c->tie; \
    const float nmax = c->nmax, pmax = c->pmax;                                                   \
    for (size_t r = r0; r < r1; ++r) {                                                            \
        const IN_T *xr = x + (size_t)r * row_len;                                                 \
        OUT_T *orow = out + (size_t)r * row_len;                                                  \
        float an = 0.0f, ap = 0.0f;                                                               \
        for (size_t i = 0; i < row_len; ++i) {                                                    \
            float f = (float)xr[i];                                                               \
            float xn = (f <= 0.0f) ? f : 0.0f, xp = (f > 0.0f) ? f : 0.0f;   /* qu.py:428-429 */  \
            if (fabsf(xn) > an) an = fabsf(xn);                                                   \
            if (xp > ap) ap = xp;                                                                 \
        }                                                                                         \
        volatile float sn = RND_IN(an / nmax);                  /* qu.py:432 */                   \
        volatile float sp = RND_IN(ap / pmax);                  /* qu.py:433 */                   \
        for (size_t i = 0; i < row_len; ++i) {                                                    \
            float f = (float)xr[i];                                                               \
            float xn = (f <= 0.0f) ? f : 0.0f, xp = (f > 0.0f) ? f : 0.0f;                        \
            volatile float vn = RND_IN(xn / sn);                /* qu.py:436 */                   \
            volatile float vp = RND_IN(xp / sp);                /* qu.py:437 */                   \
            float qn = round_rule(vn, gneg, kn, tie);           /* qu.py:443 */                   \
            float qp = round_rule(vp, gpos, kp, tie);           /* qu.py:444 */                   \
            volatile float o;                                                                     \
            if (tie == 0) {                                                                       \
                volatile float tn = qn * sn, tp = qp * sp;                                        \
                o = tn + tp;                                    /* qu.py:450 */                   \
            } else {                                                                              \
                volatile float qs = qn + qp;                                                      \
                o = qs * ((f <= 0.0f) ? sn : sp);               /* qu.py:409-410 */               \
            }                                                                                     \
            orow[i] = (OUT_T)o;                                                                   \
        }                                                                                         \
    }                                                                                             \
}                                                                                                 \
void NAME(const IN_T *x, OUT_T *out, size_t n_rows, size_t row_len, const float *gneg, int kn,    \
          float nmax, const float *gpos, int kp, float pmax, int tie)                             \
{                                                                                                 \
    NAME##_ctx c = {x, out, row_len, gneg, kn, nmax, gpos, kp, pmax, tie};                        \
    parallel_rows(NAME##_range, &c, n_rows);                                                      \
}

DEFINE_SIGNSPLIT(port_signsplit_f32_f32, float, float, RND_F32)
DEFINE_SIGNSPLIT(port_signsplit_f16_f16, f16, f16, RND_F16)
DEFINE_SIGNSPLIT(port_signsplit_f16_f32, f16, float, RND_F16)

/* ---- transform + rotate (+ quant) ----------------------------------------------------- */
/* q128: the dense 128x128 fp32 block of Q (row-major, q[i*128+j]); rotated may be NULL.
 * grid == NULL skips the quantizer (out then receives the rotated fp16 values). */
typedef struct { const float *x, *smooth, *q128; f16 *out, *rotated; size_t cpr; const float *grid; int k; float gmax; } trq_ctx;
static void trq_range(const void *vc, size_t c0, size_t c1)
{
    const trq_ctx *t = (const trq_ctx *)vc;
    const float *x = t->x, *smooth = t->smooth, *q128 = t->q128, *grid = t->grid;
    f16 *out = t->out, *rotated = t->rotated;
    const size_t cpr = t->cpr; const int k = t->k; const float gmax = t->gmax;
    for (size_t c = c0; c < c1; ++c) {
        const float *xc = x + (size_t)c * 128;
        const float *sc = smooth ? smooth + ((size_t)c % cpr) * 128 : NULL;
        float xs[128], acc[128];
        for (int i = 0; i < 128; ++i) {
            volatile float m = sc ? xc[i] * sc[i] : xc[i];     /* basic_var.py:263 `.mul(s)`, fp32 */
            xs[i] = m;
            acc[i] = 0.0f;
        }
        for (int i = 0; i < 128; ++i) {                          /* matmul(x, Q): fp32 accumulate */
            const float xi = xs[i];
            const float *qr = q128 + (size_t)i * 128;
            for (int j = 0; j < 128; ++j) acc[j] += xi * qr[j];
        }
        f16 y[128];
        float a = 0.0f;
        for (int j = 0; j < 128; ++j) {
            y[j] = (f16)acc[j];                                  /* autocast fp16 GEMM output */
            a = max_nan(a, fabsf((float)y[j]));
        }
        if (rotated) memcpy(rotated + (size_t)c * 128, y, sizeof(y));
        f16 *oc = out + (size_t)c * 128;
        if (!grid) { memcpy(oc, y, sizeof(y)); continue; }
        volatile float s = RND_F16(a / gmax);
        for (int j = 0; j < 128; ++j) {
            volatile float v = RND_F16((float)y[j] / s);
            float q = scan_rule(v, grid, k);
            volatile float o = q * s;
            oc[j] = (f16)o;
        }
    }
}

void port_transform_rotate_quant(const float *x, const float *smooth, const float *q128, f16 *out, f16 *rotated,
                                 size_t n_rows, size_t n_cols, const float *grid, int k, float gmax)
{
    trq_ctx t = {x, smooth, q128, out, rotated, n_cols / 128, grid, k, gmax};
    parallel_rows(trq_range, &t, n_rows * (n_cols / 128));
}
