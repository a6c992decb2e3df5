/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_raytrace.c header).
 *
 * CPU restatement (plain C, float32 arithmetic exactly as numpy performs it) of the
 * reference LOS sampler
 *   /root/reference/raytracingGRFF/gpu_raytrace.py:489-535  (_trilinear_numpy_uniform)
 *   /root/reference/raytracingGRFF/gpu_raytrace.py:473-486  (_compute_ds_from_valid)
 *   /root/reference/raytracingGRFF/gpu_raytrace.py:632-651  (_sample_model_with_rays_cpu)
 *
 * Parity status: PINNED — checked against the reference's own CPU sampler on the
 * reference's test fixture (tests/test_gpu_raytrace.py:13-44) and on bench_raytrace.make_case,
 * run in the build container (tests/golden/make_golden.py).
 *
 * Build with -ffp-contract=off: numpy never fuses a multiply with an add.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* One field, one sample; gpu_raytrace.py:495-534.  x0/inv_d arrive as the Python
 * floats the reference passes and are rounded to f32 by numpy's weak-scalar rule. */
static inline float trilinear_f32(const float *f, int nx, int ny, int nz,
                                  float px, float py, float pz,
                                  float x0, float y0, float z0,
                                  float inv_dx, float inv_dy, float inv_dz, float fill)
{
    const float fx = (px - x0) * inv_dx;
    const float fy = (py - y0) * inv_dy;
    const float fz = (pz - z0) * inv_dz;
    const int inb = (fx >= 0.0f) && (fy >= 0.0f) && (fz >= 0.0f) &&
                    (fx <= (float)(nx - 1)) && (fy <= (float)(ny - 1)) && (fz <= (float)(nz - 1));
    if (!inb) return fill;
    int i = (int)floorf(fx), j = (int)floorf(fy), k = (int)floorf(fz);
    i = i < 0 ? 0 : (i > nx - 2 ? nx - 2 : i);
    j = j < 0 ? 0 : (j > ny - 2 ? ny - 2 : j);
    k = k < 0 ? 0 : (k > nz - 2 ? nz - 2 : k);
    /* fx - ii is evaluated in f64 by numpy (f32 - i32 promotes) and is exact; clip; cast */
    double dtx = (double)fx - (double)i, dty = (double)fy - (double)j, dtz = (double)fz - (double)k;
    dtx = dtx < 0.0 ? 0.0 : (dtx > 1.0 ? 1.0 : dtx);
    dty = dty < 0.0 ? 0.0 : (dty > 1.0 ? 1.0 : dty);
    dtz = dtz < 0.0 ? 0.0 : (dtz > 1.0 ? 1.0 : dtz);
    const float tx = (float)dtx, ty = (float)dty, tz = (float)dtz;
    const size_t sx = (size_t)ny * nz, sy = (size_t)nz;
    const size_t o = (size_t)i * sx + (size_t)j * sy + (size_t)k;
    const float c000 = f[o], c100 = f[o + sx], c010 = f[o + sy], c110 = f[o + sx + sy];
    const float c001 = f[o + 1], c101 = f[o + sx + 1], c011 = f[o + sy + 1], c111 = f[o + sx + sy + 1];
    const float c00 = c000 * (1.0f - tx) + c100 * tx;
    const float c10 = c010 * (1.0f - tx) + c110 * tx;
    const float c01 = c001 * (1.0f - tx) + c101 * tx;
    const float c11 = c011 * (1.0f - tx) + c111 * tx;
    const float c0 = c00 * (1.0f - ty) + c10 * ty;
    const float c1 = c01 * (1.0f - ty) + c11 * ty;
    return c0 * (1.0f - tz) + c1 * tz;
}

static inline float dist_f32(const float *a, const float *b)
{
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return sqrtf((dx * dx + dy * dy) + dz * dz);
}

/* First segment, gpu_raytrace.py:482: np.linalg.norm of a 1-D float32 vector goes through BLAS
 * sdot, which (OpenBLAS x86-64, as shipped with numpy here) accumulates the float products in a
 * double before rounding to float; the axis=1 norms of :484 stay in float32. */
static inline float dist_first_f32(const float *a, const float *b)
{
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    double acc = 0.0;
    acc += (double)(dx * dx);
    acc += (double)(dy * dy);
    acc += (double)(dz * dz);
    return sqrtf((float)acc);
}

/*
 * _sample_model_with_rays_cpu.  pos (n_rec,n_rays,3) f32, s (n_rec,n_rays) f32,
 * ray_start (n_rays,3) f32, fields (nx,ny,nz) f32 C-order; outputs (n_rec,n_rays).
 * x0/dx etc. are the doubles returned by _check_uniform_grid.
 */
int oracle_sample_model(const float *ne_xyz, const float *te_xyz, const float *b_xyz,
                        int nx, int ny, int nz,
                        double x0, double dx, double y0, double dy, double z0, double dz,
                        const float *pos, const float *s, const float *ray_start,
                        long n_rec, long n_rays, double r_sun_cm,
                        double fill_ne, double fill_te, double fill_b,
                        float *ne, float *te, float *b, float *ds, uint8_t *valid)
{
    const float fx0 = (float)x0, fy0 = (float)y0, fz0 = (float)z0;
    const float idx = (float)(1.0 / dx), idy = (float)(1.0 / dy), idz = (float)(1.0 / dz);
    const float rs = (float)r_sun_cm; /* numpy 2: f32 array * python float -> f32 */
#pragma omp parallel for schedule(static)
    for (long r = 0; r < n_rays; ++r) {
        const float *prev = ray_start + r * 3;
        int first = 1;
        for (long i = 0; i < n_rec; ++i) {
            const size_t q = (size_t)i * n_rays + r;
            const float *p = pos + q * 3;
            const float sv = s[q];
            const int v = isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]) && isfinite(sv) && (sv > 0.0f);
            valid[q] = (uint8_t)v;
            ne[q] = trilinear_f32(ne_xyz, nx, ny, nz, p[0], p[1], p[2], fx0, fy0, fz0, idx, idy, idz, (float)fill_ne);
            te[q] = trilinear_f32(te_xyz, nx, ny, nz, p[0], p[1], p[2], fx0, fy0, fz0, idx, idy, idz, (float)fill_te);
            b[q] = trilinear_f32(b_xyz, nx, ny, nz, p[0], p[1], p[2], fx0, fy0, fz0, idx, idy, idz, (float)fill_b);
            if (v) {
                ds[q] = (first ? dist_first_f32(p, prev) : dist_f32(p, prev)) * rs;
                prev = p;
                first = 0;
            } else {
                ds[q] = 0.0f;
            }
        }
    }
    return 0;
}
