/*
 * dog_oracle.c — CPU restatement of PawsomeTracker.jl's Tracker hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * timed CPU baseline.  The product path is libpawsome_cuda.so and has no CPU
 * fallback.
 *
 * PARITY UNPINNED.  The arithmetic of the reference lives in the third-party
 * package ImageFiltering.jl (Project.toml:33, compat "0.4, 0.5, 0.6, 0.7",
 * no Manifest => unpinned) which is not vendored under /root/reference, and
 * neither julia nor ffmpeg exist in this image, so the reference cannot be
 * executed and its tests hold no golden vector for this boundary
 * (test/test-basic-test.jl:139-148 asserts nothing about positions).  This
 * file restates the published algorithm of
 *   - Kernel.DoG / KernelFactors.gaussian  (ImageFiltering.jl, Kernel module)
 *   - imfilter!(…FIR…, NoPad(), inds)      (ImageFiltering.jl, dense direct loop)
 *   - PaddedViews.PaddedView               (constant fill outside the frame)
 *   - StatsBase.mode                       (first value to reach the max count)
 *   - Base.findmax                         (first maximum, column-major order)
 * anchored on the reference's own call sites cited at each function.
 *
 * Conventions: frames are row-major uint8, H rows, W columns, `pitch` bytes
 * between rows (the memory layout of the reference's
 * PermutedDimsArray{Gray{N0f8},2,(2,1)} over a W×H Matrix,
 * src/PawsomeTracker.jl:36).  Indices crossing this API are 1-based
 * (row, col) like the reference's CartesianIndex; internals are 0-based.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction so
 * the dense loop's rounding sequence is the reference's tmp += a*k).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define PTO_API __attribute__((visibility("default")))

/* get_sigma: src/PawsomeTracker.jl:30  (target_width / 2sqrt(2log(2))) */
PTO_API double pto_sigma(double target_width)
{
    return target_width / (2.0 * sqrt(2.0 * log(2.0)));
}

/* Kernel.DoG(σ) length: the wide Gaussian σm = σ·√2 gets the default length
 * 4⌈σm⌉+1 and the narrow one is built at the same length
 * (ImageFiltering Kernel.DoG(σps) ; call site src/PawsomeTracker.jl:43). */
PTO_API int pto_kernel_len(double target_width)
{
    double sm = pto_sigma(target_width) * sqrt(2.0);
    return 4 * (int)ceil(sm) + 1;
}

/* guess_window_size: src/PawsomeTracker.jl:64-68 (uses the NARROW σ) */
PTO_API int pto_default_window(double target_width)
{
    return 4 * (int)ceil(pto_sigma(target_width)) + 1;
}

/* KernelFactors.gaussian(σ, l): g[x] = exp(-x²/(2σ²)), x = -w..w, divided by
 * its own discrete sum over exactly l taps. */
static void gaussian_factor(double sigma, int l, double *g)
{
    int w = l >> 1;
    double s = 0.0;
    for (int k = 0; k < l; ++k) {
        double x = (double)(k - w);
        g[k] = exp(-(x * x) / (2.0 * sigma * sigma));
        s += g[k];
    }
    for (int k = 0; k < l; ++k) g[k] /= s;
}

/* The two 1-D factors of the DoG: gp (σ) and gm (σ√2), each length l. */
PTO_API int pto_factors(double target_width, double *gp, double *gm)
{
    int l = pto_kernel_len(target_width);
    double s = pto_sigma(target_width);
    gaussian_factor(s, l, gp);
    gaussian_factor(s * sqrt(2.0), l, gm);
    return l;
}

/* kernel = direction * Kernel.DoG(σ): src/PawsomeTracker.jl:42-43.
 * K is returned kernel-column-major: K[b*l + a], a = kernel row offset index,
 * b = kernel column offset index (the storage order of the Julia matrix). */
PTO_API int pto_dense_kernel(double target_width, int darker, double *K)
{
    int l = pto_kernel_len(target_width);
    double *gp = (double *)malloc(sizeof(double) * 2 * (size_t)l);
    double *gm = gp + l;
    pto_factors(target_width, gp, gm);
    double dir = darker ? -1.0 : 1.0;
    for (int b = 0; b < l; ++b)
        for (int a = 0; a < l; ++a)
            K[(size_t)b * l + a] = dir * (gp[a] * gp[b] - gm[a] * gm[b]);
    free(gp);
    return l;
}

/* StatsBase.mode over the H×W view (src/PawsomeTracker.jl:47): iterate with
 * the row index fastest (column-major over the view), return the value whose
 * count is the first to reach the final maximum. */
PTO_API int pto_mode_u8(const uint8_t *frame, int H, int W, size_t pitch)
{
    long cnt[256];
    memset(cnt, 0, sizeof cnt);
    long mc = 0;
    int mv = frame[0];
    for (int x = 0; x < W; ++x)
        for (int y = 0; y < H; ++y) {
            int v = frame[(size_t)y * pitch + x];
            long c = ++cnt[v];
            if (c > mc) { mc = c; mv = v; }
        }
    return mv;
}

typedef struct {
    int i, j;          /* clamped 1-based (row, col): what trckr(guess) returns, :61 */
    int raw_i, raw_j;  /* unclamped 1-based argmax position (may lie outside the frame) */
    double resp;       /* maximum response (the value findmax discards, :59) */
    double second;     /* best response at any OTHER position (for near-tie flagging) */
    double maxabs;     /* max |R| over the window (tolerance scale) */
} pto_result;

/* Pixel of the PaddedView (src/PawsomeTracker.jl:48): N0f8 value inside the
 * frame, the fill value outside.  y, x are 0-based and may be out of range. */
static inline double padded_px(const uint8_t *f, int H, int W, size_t pitch, int fill, int y, int x)
{
    if (y < 0 || y >= H || x < 0 || x >= W) return (double)fill / 255.0;
    return (double)f[(size_t)y * pitch + x] / 255.0;
}

/* findmax over a column-major window (src/PawsomeTracker.jl:58-59): first
 * maximum with the column as the slow index; then :60-61 (absolute + clamp).
 * R is window-column-major: R[xx*wr + yy]. y0/x0 = 0-based frame coords of the
 * window origin. */
static void finish_argmax(const double *R, int wr, int wc, int y0, int x0, int H, int W, pto_result *out)
{
    double best = R[0], second = -INFINITY, maxabs = 0.0;
    long bi = 0;
    long n = (long)wr * wc;
    for (long t = 0; t < n; ++t) {
        double v = R[t];
        double av = fabs(v);
        if (av > maxabs) maxabs = av;
        if (t == 0) continue;
        if (v > best) { second = best; best = v; bi = t; }
        else if (v > second) second = v;
    }
    int xx = (int)(bi / wr), yy = (int)(bi % wr);
    int ry = y0 + yy + 1, rx = x0 + xx + 1; /* 1-based */
    out->raw_i = ry; out->raw_j = rx;
    out->i = ry < 1 ? 1 : (ry > H ? H : ry);
    out->j = rx < 1 ? 1 : (rx > W ? W : rx);
    out->resp = best; out->second = second; out->maxabs = maxabs;
}

/* Column-major padded copy of the footprint as double: Pd[(x)*fr + y]. */
static double *footprint_colmajor(const uint8_t *f, int H, int W, size_t pitch, int fill,
                                  int fy0, int fx0, int fr, int fc)
{
    double *Pd = (double *)malloc(sizeof(double) * (size_t)fr * fc);
    if (!Pd) return NULL;
    for (int x = 0; x < fc; ++x)
        for (int y = 0; y < fr; ++y)
            Pd[(size_t)x * fr + y] = padded_px(f, H, W, pitch, fill, fy0 + y, fx0 + x);
    return Pd;
}

/*
 * Dense Float64 correlation over an output rectangle, in the reference's
 * summation order (ImageFiltering's direct FIR loop behind
 * src/PawsomeTracker.jl:57): for each output, tmp = 0; kernel column b outer,
 * kernel row a inner; tmp += P*K sequentially, no FMA contraction.
 * Eight neighbouring outputs are advanced together (independent accumulators,
 * each in the exact scalar order) so the compiler can use SIMD lanes without
 * changing any rounding.
 *
 * Rectangle: output rows y0..y0+wr-1, cols x0..x0+wc-1 (0-based frame coords,
 * may extend outside the frame).  Rout (optional) receives the response map,
 * window-column-major.
 */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void dense_rect_core(const double *Pd, int fr, const double *K, int l,
                            int wr, int wc, double *R)
{
    enum { LANES = 8 };
    for (int xx = 0; xx < wc; ++xx) {
        int yy = 0;
        for (; yy + LANES <= wr; yy += LANES) {
            double tmp[LANES];
            for (int t = 0; t < LANES; ++t) tmp[t] = 0.0;
            for (int b = 0; b < l; ++b) {
                const double *col = Pd + (size_t)(xx + b) * fr + yy;
                const double *kb = K + (size_t)b * l;
                for (int a = 0; a < l; ++a) {
                    double k = kb[a];
                    for (int t = 0; t < LANES; ++t) tmp[t] += col[a + t] * k;
                }
            }
            for (int t = 0; t < LANES; ++t) R[(size_t)xx * wr + yy + t] = tmp[t];
        }
        for (; yy < wr; ++yy) {
            double tmp = 0.0;
            for (int b = 0; b < l; ++b) {
                const double *col = Pd + (size_t)(xx + b) * fr + yy;
                const double *kb = K + (size_t)b * l;
                for (int a = 0; a < l; ++a) tmp += col[a] * kb[a];
            }
            R[(size_t)xx * wr + yy] = tmp;
        }
    }
}

PTO_API int pto_rect_dense(const uint8_t *frame, int H, int W, size_t pitch, int fill,
                           double target_width, int darker,
                           int y0, int x0, int wr, int wc,
                           pto_result *out, double *Rout)
{
    int l = pto_kernel_len(target_width);
    int w = l >> 1;
    int fr = wr + 2 * w, fc = wc + 2 * w;
    double *K = (double *)malloc(sizeof(double) * (size_t)l * l);
    double *Pd = footprint_colmajor(frame, H, W, pitch, fill, y0 - w, x0 - w, fr, fc);
    double *R = Rout ? Rout : (double *)malloc(sizeof(double) * (size_t)wr * wc);
    if (!K || !Pd || !R) { free(K); free(Pd); if (!Rout) free(R); return -1; }
    pto_dense_kernel(target_width, darker, K);
    dense_rect_core(Pd, fr, K, l, wr, wc, R);
    finish_argmax(R, wr, wc, y0, x0, H, W, out);
    free(K); free(Pd); if (!Rout) free(R);
    return 0;
}

/*
 * Separable Float64 evaluation of the same response (row pass with both
 * Gaussians, then column pass, subtract, sign).  Mathematically identical to
 * the dense sum; rounding differs at the 1e-16 level.  Used where the dense
 * loop would take minutes (4K / tw=100, full-frame), after being validated
 * against pto_rect_dense on small shapes (tests/test_oracle.py).
 */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void separable_core(const double *Pd, int fr, int fc, const double *gp, const double *gm,
                           int l, double dir, int wr, int wc, double *R)
{
    /* horizontal (over x) pass: for every footprint row y and output col xx */
    double *Mp = (double *)malloc(sizeof(double) * (size_t)fr * wc * 2);
    double *Mm = Mp + (size_t)fr * wc;
    (void)fc;
    for (int xx = 0; xx < wc; ++xx) {
        double *mp = Mp + (size_t)xx * fr, *mm = Mm + (size_t)xx * fr;
        for (int y = 0; y < fr; ++y) { mp[y] = 0.0; mm[y] = 0.0; }
        for (int b = 0; b < l; ++b) {
            const double *col = Pd + (size_t)(xx + b) * fr;
            double kp = gp[b], km = gm[b];
            for (int y = 0; y < fr; ++y) { mp[y] += col[y] * kp; mm[y] += col[y] * km; }
        }
    }
    for (int xx = 0; xx < wc; ++xx) {
        const double *mp = Mp + (size_t)xx * fr, *mm = Mm + (size_t)xx * fr;
        for (int yy = 0; yy < wr; ++yy) {
            double sp = 0.0, sm = 0.0;
            for (int a = 0; a < l; ++a) { sp += mp[yy + a] * gp[a]; sm += mm[yy + a] * gm[a]; }
            R[(size_t)xx * wr + yy] = dir * (sp - sm);
        }
    }
    free(Mp);
}

PTO_API int pto_rect_separable(const uint8_t *frame, int H, int W, size_t pitch, int fill,
                               double target_width, int darker,
                               int y0, int x0, int wr, int wc,
                               pto_result *out, double *Rout)
{
    int l = pto_kernel_len(target_width);
    int w = l >> 1;
    int fr = wr + 2 * w, fc = wc + 2 * w;
    double *g = (double *)malloc(sizeof(double) * 2 * (size_t)l);
    double *Pd = footprint_colmajor(frame, H, W, pitch, fill, y0 - w, x0 - w, fr, fc);
    double *R = Rout ? Rout : (double *)malloc(sizeof(double) * (size_t)wr * wc);
    if (!g || !Pd || !R) { free(g); free(Pd); if (!Rout) free(R); return -1; }
    pto_factors(target_width, g, g + l);
    separable_core(Pd, fr, fc, g, g + l, l, darker ? -1.0 : 1.0, wr, wc, R);
    finish_argmax(R, wr, wc, y0, x0, H, W, out);
    free(g); free(Pd); if (!Rout) free(R);
    return 0;
}

/*
 * (trckr::Tracker)(guess): src/PawsomeTracker.jl:55-62.
 * guess (gi, gj) is 1-based (row, col); radii = window_size .÷ 2 (:44).
 * window_indices = guess .- radii : guess .+ radii (:56).
 * dense != 0 → reference-order dense loop; else separable.
 */
PTO_API int pto_tracker_step(const uint8_t *frame, int H, int W, size_t pitch, int fill,
                             double target_width, int darker, int ws_rows, int ws_cols,
                             int gi, int gj, int dense, pto_result *out, double *Rout)
{
    int rr = ws_rows / 2, rc = ws_cols / 2;
    int y0 = (gi - 1) - rr, x0 = (gj - 1) - rc;
    int wr = 2 * rr + 1, wc = 2 * rc + 1;
    return dense ? pto_rect_dense(frame, H, W, pitch, fill, target_width, darker, y0, x0, wr, wc, out, Rout)
                 : pto_rect_separable(frame, H, W, pitch, fill, target_width, darker, y0, x0, wr, wc, out, Rout);
}

/*
 * One lock-step time step over n independent videos (the reference would run
 * n separate track() calls; each does src/PawsomeTracker.jl:166-167 once per
 * frame).  Dense reference-order loop per video, OpenMP across videos: the
 * "multi-threaded CPU" baseline figure of BASELINE.md §3.  guess/out are
 * n×2 int32 (1-based row, col).  Returns threads used.
 */
typedef struct {
    void (*fn)(void *ctx, int idx);
    void *ctx;
    int n;
    volatile int next;
} pfor_t;

static void *pfor_worker(void *arg)
{
    pfor_t *p = (pfor_t *)arg;
    for (;;) {
        int i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (i >= p->n) break;
        p->fn(p->ctx, i);
    }
    return NULL;
}

PTO_API int pto_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* Dynamic parallel-for over n items on nthreads POSIX threads (0 = all cores). */
static int parallel_for(int n, int nthreads, void (*fn)(void *, int), void *ctx)
{
    if (nthreads <= 0) nthreads = pto_max_threads();
    if (nthreads > n) nthreads = n;
    if (nthreads < 1) nthreads = 1;
    pfor_t p = { fn, ctx, n, 0 };
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&th[started], NULL, pfor_worker, &p) == 0) ++started;
    pfor_worker(&p);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return started + 1;
}

typedef struct {
    const uint8_t *const *frames; int H, W; size_t pitch; const int *fills;
    double tw; int darker, ws_rows, ws_cols; const int *guess; int *out_ij; double *out_resp;
} batch_ctx;

static void batch_item(void *c, int v)
{
    batch_ctx *b = (batch_ctx *)c;
    pto_result r;
    pto_tracker_step(b->frames[v], b->H, b->W, b->pitch, b->fills[v], b->tw, b->darker,
                     b->ws_rows, b->ws_cols, b->guess[2 * v], b->guess[2 * v + 1], 1, &r, NULL);
    b->out_ij[2 * v] = r.i; b->out_ij[2 * v + 1] = r.j;
    if (b->out_resp) b->out_resp[v] = r.resp;
}

PTO_API int pto_batch_step_dense(const uint8_t *const *frames, int n, int H, int W, size_t pitch,
                                 const int *fills, double target_width, int darker,
                                 int ws_rows, int ws_cols, const int *guess, int *out_ij,
                                 double *out_resp, int nthreads)
{
    batch_ctx b = { frames, H, W, pitch, fills, target_width, darker, ws_rows, ws_cols,
                    guess, out_ij, out_resp };
    return parallel_for(n, nthreads, batch_item, &b);
}

/* Same, on one rectangle per call, threads across output columns: used to time
 * the dense full-frame / auto-detect pass with all host cores. */
typedef struct { const double *Pd; int fr; const double *K; int l, wr, wc, chunk; double *R; } rect_ctx;

static void rect_item(void *c, int idx)
{
    rect_ctx *r = (rect_ctx *)c;
    int c0 = idx * r->chunk;
    int n = r->wc - c0 < r->chunk ? r->wc - c0 : r->chunk;
    dense_rect_core(r->Pd + (size_t)c0 * r->fr, r->fr, r->K, r->l, r->wr, n, r->R + (size_t)c0 * r->wr);
}

PTO_API int pto_rect_dense_mt(const uint8_t *frame, int H, int W, size_t pitch, int fill,
                              double target_width, int darker,
                              int y0, int x0, int wr, int wc, int nthreads,
                              pto_result *out)
{
    int l = pto_kernel_len(target_width);
    int w = l >> 1;
    int fr = wr + 2 * w, fc = wc + 2 * w;
    double *K = (double *)malloc(sizeof(double) * (size_t)l * l);
    double *Pd = footprint_colmajor(frame, H, W, pitch, fill, y0 - w, x0 - w, fr, fc);
    double *R = (double *)malloc(sizeof(double) * (size_t)wr * wc);
    if (!K || !Pd || !R) { free(K); free(Pd); free(R); return -1; }
    pto_dense_kernel(target_width, darker, K);
    rect_ctx rc = { Pd, fr, K, l, wr, wc, 4, R };
    int used = parallel_for((wc + rc.chunk - 1) / rc.chunk, nthreads, rect_item, &rc);
    finish_argmax(R, wr, wc, y0, x0, H, W, out);
    free(K); free(Pd); free(R);
    return used;
}
