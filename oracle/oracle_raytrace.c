/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (raytracinggrff_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement (plain C, float64) of the reference ray integrator
 *   /root/reference/raytracingGRFF/build_rays.py:128-248  (ray_trace)
 * including the scipy RegularGridInterpolator(linear, bounds_error=False,
 * fill_value=nan) semantics it relies on (build_rays.py:140-143) and numpy's
 * np.gradient (build_rays.py:136-138).
 *
 * Parity status: PINNED — checked against the reference's own ray_trace run in
 * the build container (tests/golden/make_golden.py -> tests/golden/*.npz,
 * tests/test_oracle_golden.py).
 *
 * The reference vectorises every step over all rays; rays never interact, so
 * this restatement loops per ray (OpenMP over rays) with the same per-ray
 * arithmetic, in the same evaluation order as the numpy expressions.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* build_rays.py:29-32 */
#define ORACLE_C_R (2.998e10 / 6.96e10)

typedef struct {
    const double *w, *gx, *gy, *gz; /* (nx,ny,nz) C-order: x slowest, z fastest */
    const double *xg, *yg, *zg;
    int nx, ny, nz;
} cube_t;

/* np.gradient(f, h, axis) with uniform scalar spacing, edge_order=1
 * (build_rays.py:136-138): interior (f[i+1]-f[i-1])/(2h), faces one-sided /h. */
void oracle_gradient(const double *f, int nx, int ny, int nz, double h, int axis, double *out)
{
    const int n[3] = {nx, ny, nz};
    const size_t st[3] = {(size_t)ny * nz, (size_t)nz, 1};
    const int na = n[axis];
    const size_t sa = st[axis];
    const double h2 = 2.0 * h;
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                const size_t o = (size_t)i * st[0] + (size_t)j * st[1] + (size_t)k;
                const int a = axis == 0 ? i : (axis == 1 ? j : k);
                double v;
                if (na < 2)
                    v = 0.0;
                else if (a == 0)
                    v = (f[o + sa] - f[o]) / h;
                else if (a == na - 1)
                    v = (f[o] - f[o - sa]) / h;
                else
                    v = (f[o + sa] - f[o - sa]) / h2;
                out[o] = v;
            }
}

/* scipy find_interval_ascending + clip to [0, n-2] for an ascending grid.
 * Starts from the uniform-grid guess and walks to the interval that satisfies
 * g[i] <= x < g[i+1] (last interval closed on the right). */
static inline int find_cell(const double *g, int n, double x)
{
    const double h = g[1] - g[0];
    int i = (int)floor((x - g[0]) / h);
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i > 0 && x < g[i]) --i;
    while (i < n - 2 && x >= g[i + 1]) ++i;
    return i;
}

/* Four scipy linear interpolators evaluated at one point (build_rays.py:140-143,
 * scipy _rgi.py _evaluate_linear): weighted 8-corner sum in itertools.product
 * order, weight = ((1*wx)*wy)*wz, NaN outside [g0, g_last] or for NaN input. */
static inline void interp4(const cube_t *c, double x, double y, double z, int want_grad,
                           double *w, double *gx, double *gy, double *gz)
{
    const double nan = NAN;
    if (isnan(x) || isnan(y) || isnan(z) || x < c->xg[0] || x > c->xg[c->nx - 1] ||
        y < c->yg[0] || y > c->yg[c->ny - 1] || z < c->zg[0] || z > c->zg[c->nz - 1]) {
        *w = nan;
        if (want_grad) { *gx = nan; *gy = nan; *gz = nan; }
        return;
    }
    const int i = find_cell(c->xg, c->nx, x);
    const int j = find_cell(c->yg, c->ny, y);
    const int k = find_cell(c->zg, c->nz, z);
    const double tx = (x - c->xg[i]) / (c->xg[i + 1] - c->xg[i]);
    const double ty = (y - c->yg[j]) / (c->yg[j + 1] - c->yg[j]);
    const double tz = (z - c->zg[k]) / (c->zg[k + 1] - c->zg[k]);
    const double wx[2] = {1.0 - tx, tx}, wy[2] = {1.0 - ty, ty}, wz[2] = {1.0 - tz, tz};
    const size_t sx = (size_t)c->ny * c->nz, sy = (size_t)c->nz;
    double a = 0.0, b = 0.0, cc = 0.0, d = 0.0;
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                const double wt = ((1.0 * wx[di]) * wy[dj]) * wz[dk];
                const size_t o = (size_t)(i + di) * sx + (size_t)(j + dj) * sy + (size_t)(k + dk);
                a = a + c->w[o] * wt;
                if (want_grad) {
                    b = b + c->gx[o] * wt;
                    cc = cc + c->gy[o] * wt;
                    d = d + c->gz[o] * wt;
                }
            }
    *w = a;
    if (want_grad) { *gx = b; *gy = cc; *gz = d; }
}

/* rhs, build_rays.py:158-175. s = [r(3), k(3)] -> ds/dt. */
static inline void rhs(const cube_t *c, const double s[6], double out[6])
{
    double w, gx, gy, gz;
    interp4(c, s[0], s[1], s[2], 1, &w, &gx, &gy, &gz);
    const double k2 = (s[3] * s[3] + s[4] * s[4]) + s[5] * s[5];
    const double om = sqrt(w * w + k2);
    const int valid = isfinite(w) && isfinite(om) && (om > 0.0);
    if (!valid) {
        for (int m = 0; m < 6; ++m) out[m] = 0.0;
        return;
    }
    const double cr_om = ORACLE_C_R / om;        /* C_R / omega            */
    const double a = -w / om;                    /* -omega_pe / omega      */
    out[0] = cr_om * s[3];
    out[1] = cr_om * s[4];
    out[2] = cr_om * s[5];
    out[3] = a * gx * ORACLE_C_R;
    out[4] = a * gy * ORACLE_C_R;
    out[5] = a * gz * ORACLE_C_R;
}

/* rk4_step, build_rays.py:177-182. */
static inline void rk4_step(const cube_t *c, const double s[6], double dt, double out[6])
{
    double k1[6], k2[6], k3[6], k4[6], t[6];
    const double hdt = 0.5 * dt, c6 = dt / 6.0;
    rhs(c, s, k1);
    for (int m = 0; m < 6; ++m) t[m] = s[m] + hdt * k1[m];
    rhs(c, t, k2);
    for (int m = 0; m < 6; ++m) t[m] = s[m] + hdt * k2[m];
    rhs(c, t, k3);
    for (int m = 0; m < 6; ++m) t[m] = s[m] + dt * k3[m];
    rhs(c, t, k4);
    for (int m = 0; m < 6; ++m)
        out[m] = s[m] + c6 * (((k1[m] + 2.0 * k2[m]) + 2.0 * k3[m]) + k4[m]);
}

static inline void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

static inline double norm3(const double a[3])
{
    return sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
}

/* One step's cross-section ratio, build_rays.py:209-239 (basis :188-201). */
static inline double cross_section_ratio(const cube_t *c, const double s0[6], const double s1[6],
                                         double dt, double perturb_ratio)
{
    double rd[3], th[3], a[3] = {0, 0, 0}, e1[3], e2[3];
    for (int m = 0; m < 3; ++m) rd[m] = s1[m] - s0[m];
    const double nrd = norm3(rd);
    for (int m = 0; m < 3; ++m) th[m] = rd[m] / (nrd + 1e-32);
    if (fabs(th[2]) < 0.9) a[2] = 1.0; else a[1] = 1.0;
    cross3(a, th, e1);
    const double n1 = norm3(e1) + 1e-30;
    for (int m = 0; m < 3; ++m) e1[m] /= n1;
    cross3(th, e1, e2);
    const double n2 = norm3(e2) + 1e-30;
    for (int m = 0; m < 3; ++m) e2[m] /= n2;
    const double eps = perturb_ratio * nrd;
    double p1[6], p2[6], q1[6], q2[6], d1[3], d2[3], cr[3];
    for (int m = 0; m < 3; ++m) {
        p1[m] = s0[m] + eps * e1[m];
        p2[m] = s0[m] + eps * e2[m];
        p1[m + 3] = s0[m + 3];
        p2[m + 3] = s0[m + 3];
    }
    rk4_step(c, p1, dt, q1);
    rk4_step(c, p2, dt, q2);
    for (int m = 0; m < 3; ++m) { d1[m] = q1[m] - s1[m]; d2[m] = q2[m] - s1[m]; }
    cross3(d1, d2, cr);
    const double dot = (cr[0] * th[0] + cr[1] * th[1]) + cr[2] * th[2];
    return fabs(dot) / (eps * eps);
}

/*
 * ray_trace, build_rays.py:128-248.
 *  omega_pe (nx,ny,nz) C-order f64; grids f64; starts (n_rays,) f64; kvec (n_rays,3) f64;
 *  r_record out (n_rec, n_rays, 3), s_record out (n_rec, n_rays) (untouched if !trace_cs),
 *  n_rec = ceil(n_steps / record_stride); record taken AFTER step i when i % stride == 0.
 *  active_steps (optional): number of central-ray steps taken while the ray could still
 *  move (bookkeeping for the benchmark, not part of the reference).
 * Returns 0, or -1 on allocation failure.
 */
/* Thread count of every OpenMP region of the oracle (ray_trace, sampler, get_mw_slice): a launcher such as
 * torchrun exports OMP_NUM_THREADS=1 to its workers, which would serialise the CPU baseline. */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_ray_trace(const double *omega_pe, const double *xg, const double *yg, const double *zg,
                     int nx, int ny, int nz, double freq_hz,
                     const double *x_start, const double *y_start, const double *z_start,
                     const double *kvec, long n_rays, double dt, long n_steps, long record_stride,
                     int trace_cs, double perturb_ratio, int n_threads,
                     double *r_record, double *s_record, long long *active_steps)
{
    /* n_threads < 0: "keep the gradient of this cube between calls", thread count -1 - n_threads (0 = leave).
     * The reference recomputes np.gradient in every ray_trace call (build_rays.py:136-138); over a full map
     * that is amortised over millions of rays, over the benchmark's sub-samples of a 512^3 cube it would dominate. */
    static const double *cache_key = NULL;
    static double *cache_g[3] = {NULL, NULL, NULL};
    static int cache_dims[3] = {0, 0, 0};
    static double cache_h[3] = {0, 0, 0};
    static double cache_sum = 0.0;
    const int keep = n_threads < 0;
    if (keep) n_threads = -1 - n_threads;
    const size_t nvox = (size_t)nx * ny * nz;
    const double hx = xg[1] - xg[0], hy = yg[1] - yg[0], hz = zg[1] - zg[0];
    double *gx, *gy, *gz;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    double csum = 0.0;                       /* a few samples of the content: the address alone can be recycled */
    for (size_t q = 0; q < 64; ++q) csum += omega_pe[(nvox - 1) / 63 * q] * (double)(q + 1);
    const int hit = keep && cache_key == omega_pe && cache_sum == csum && cache_dims[0] == nx && cache_dims[1] == ny && cache_dims[2] == nz &&
                    cache_h[0] == hx && cache_h[1] == hy && cache_h[2] == hz;
    if (hit) {
        gx = cache_g[0]; gy = cache_g[1]; gz = cache_g[2];
    } else {
        if (cache_key) { free(cache_g[0]); free(cache_g[1]); free(cache_g[2]); cache_key = NULL; }
        gx = (double *)malloc(nvox * sizeof(double));
        gy = (double *)malloc(nvox * sizeof(double));
        gz = (double *)malloc(nvox * sizeof(double));
        if (!gx || !gy || !gz) { free(gx); free(gy); free(gz); return -1; }
        oracle_gradient(omega_pe, nx, ny, nz, hx, 0, gx);
        oracle_gradient(omega_pe, nx, ny, nz, hy, 1, gy);
        oracle_gradient(omega_pe, nx, ny, nz, hz, 2, gz);
        if (keep) {
            cache_key = omega_pe; cache_g[0] = gx; cache_g[1] = gy; cache_g[2] = gz;
            cache_dims[0] = nx; cache_dims[1] = ny; cache_dims[2] = nz;
            cache_h[0] = hx; cache_h[1] = hy; cache_h[2] = hz;
            cache_sum = csum;
        }
    }
    cube_t c = {omega_pe, gx, gy, gz, xg, yg, zg, nx, ny, nz};
    const double omega0 = 2.0 * M_PI * freq_hz;
    long long active = 0;

#pragma omp parallel for schedule(dynamic, 1) reduction(+ : active)
    for (long r = 0; r < n_rays; ++r) {
        double s[6], s0[6], w, d0, d1, d2;
        s[0] = x_start[r]; s[1] = y_start[r]; s[2] = z_start[r];
        interp4(&c, s[0], s[1], s[2], 0, &w, &d0, &d1, &d2);
        /* np.maximum propagates NaN: NaN start -> NaN k (build_rays.py:148-151) */
        const double arg = omega0 * omega0 - w * w;
        const double kc0 = isnan(arg) ? NAN : sqrt(arg > 0.0 ? arg : 0.0);
        for (int m = 0; m < 3; ++m) s[3 + m] = kvec[r * 3 + m] * kc0;
        double s_ratio = 0.0;
        long rec = 0;
        for (long i = 0; i < n_steps; ++i) {
            memcpy(s0, s, sizeof(s));
            rk4_step(&c, s0, dt, s);
            if (s[0] != s0[0] || s[1] != s0[1] || s[2] != s0[2]) ++active;
            if (trace_cs) s_ratio = cross_section_ratio(&c, s0, s, dt, perturb_ratio);
            if (i % record_stride == 0) {
                double *o = r_record + ((size_t)rec * n_rays + r) * 3;
                o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
                if (trace_cs) s_record[(size_t)rec * n_rays + r] = s_ratio;
                ++rec;
            }
        }
    }
    if (active_steps) *active_steps = active;
    if (!keep) { free(gx); free(gy); free(gz); }
    return 0;
}
