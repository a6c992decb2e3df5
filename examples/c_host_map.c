/*
 * A host written in plain C against include/rtgrff.h — no Python, no torch: what a maintainer binding the library
 * from another language sees.  It builds a small analytic corona in host memory, uploads it, shards the image rows
 * as a rank of a (here single-rank) job, renders its share with the fused kernel into a device slab, gathers the
 * image and prints it as text, so that tests/test_c_host.py can compare it number for number with the same map
 * rendered through the Python API.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_host_map.c -o build/c_host_map \
 *       -Lraytracinggrff_b200 -lrtgrff_b200 -Wl,-rpath,$PWD/raytracinggrff_b200 -lm
 *   build/c_host_map 16 48 120e6 > map.txt        # n_pix grid_n freq_hz
 *
 * In a multi-process job every rank does the same with its own (world, rank): rank 0 calls rtgrff_comm_unique_id and
 * ships the 128 bytes to the others (MPI_Bcast, a file, a socket), all call rtgrff_comm_init_rank, and
 * rtgrff_gather_image becomes a collective.  The reference's counterpart is the ProcessPoolExecutor chunking of
 * script/resample_with_ray_tracing.py:333-352.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtgrff.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != RTGRFF_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, rtgrff_last_error());        \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char **argv)
{
    const int n_pix = argc > 1 ? atoi(argv[1]) : 16, n = argc > 2 ? atoi(argv[2]) : 48;
    const double freq = argc > 3 ? atof(argv[3]) : 120e6;
    const double extent = 3.0, x_fov = 1.3, z_obs = 3.0, r_sun_cm = 6.957e10, pi = 3.14159265358979323846;
    const size_t nvox = (size_t)n * n * n;

    /* --- the model: n_e = 4.2e4 10^(4.32/r), T = 1e6 + 4e5 tanh(r-1), |B| = 2/r^3; inside r < 1: 0, 1e4, 0
     *     (the fills of script/resample_with_ray_tracing.py:269-284); omega_pe = 2 pi 8.93e3 sqrt(n_e) (:271) --- */
    double *grid = malloc(n * sizeof(double)), *omega = malloc(nvox * sizeof(double));
    float *ne = malloc(nvox * sizeof(float)), *te = malloc(nvox * sizeof(float)), *bb = malloc(nvox * sizeof(float));
    if (!grid || !omega || !ne || !te || !bb) return 2;
    for (int i = 0; i < n; ++i) grid[i] = -extent + 2.0 * extent * i / (n - 1);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k) {
                const size_t q = ((size_t)i * n + j) * n + k;
                const double r = sqrt(grid[i] * grid[i] + grid[j] * grid[j] + grid[k] * grid[k]);
                const double d = r < 0.999999 ? 0.0 : 4.2e4 * pow(10.0, 4.32 / r);
                omega[q] = 8.93e3 * sqrt(d) * 2.0 * pi;
                ne[q] = (float)d;
                te[q] = r < 0.999999 ? 1.0e4f : (float)(1.0e6 + 0.4e6 * tanh(r - 1.0));
                bb[q] = r < 0.999999 ? 0.0f : (float)(2.0 / (r * r * r));
            }
    const double step = grid[1] - grid[0];
    const double geom[12] = {grid[0], 2.0 * extent / (n - 1), grid[n - 1], step, grid[0], 2.0 * extent / (n - 1), grid[n - 1], step,
                             grid[0], 2.0 * extent / (n - 1), grid[n - 1], step};

    if (rtgrff_device_count() <= 0) {
        fprintf(stderr, "no CUDA device: %s\n", rtgrff_last_error());
        return 3;
    }
    rtgrff_ctx *ctx = NULL;
    CHECK(rtgrff_ctx_create(0, NULL, &ctx));
    CHECK(rtgrff_set_omega_cube(ctx, omega, n, n, n, geom, 0));
    CHECK(rtgrff_set_field_cubes(ctx, ne, te, bb, NULL, NULL, NULL, n, n, n, geom));

    /* --- this rank's share of the image (world 1: everything), script/resample_with_ray_tracing.py:295-303 --- */
    const int world = 1, rank = 0;
    CHECK(rtgrff_comm_init_rank(ctx, world, rank, NULL));
    int32_t *rows = malloc(n_pix * sizeof(int32_t));
    int n_local = 0, max_rows = 0;
    CHECK(rtgrff_shard_rows(n_pix, world, rank, rows, &n_local, &max_rows));
    const int64_t n_rays = (int64_t)n_local * n_pix;
    double *xs = malloc(n_rays * sizeof(double)), *ys = malloc(n_rays * sizeof(double)), *zs = malloc(n_rays * sizeof(double));
    for (int a = 0; a < n_local; ++a)
        for (int j = 0; j < n_pix; ++j) {
            const double x = -x_fov + 2.0 * x_fov * j / (n_pix - 1), y = -x_fov + 2.0 * x_fov * rows[a] / (n_pix - 1);
            xs[(size_t)a * n_pix + j] = x;
            ys[(size_t)a * n_pix + j] = y;
            zs[(size_t)a * n_pix + j] = sqrt(fabs(4.0 * z_obs * z_obs - x * x - y * y)) / 2.0;
        }

    /* --- render into a device slab [tb | vi][max_rows][n_pix], gather, print --- */
    const size_t plane = (size_t)max_rows * n_pix;
    void *slab = NULL;
    CHECK(rtgrff_device_alloc(ctx, &slab, 2 * plane * sizeof(double)));
    rtgrff_freq_params fp = {freq, 6e-3 * sqrt(100e6 / freq), 3000, 6};
    const double pix = 2.0 * x_fov / n_pix * r_sun_cm;
    int64_t stats[4];
    CHECK(rtgrff_render_map(ctx, n_rays, xs, ys, zs, NULL, NULL, 1, &fp, 1, 2.0, pix * pix, r_sun_cm, 5, 30, 0, RTGRFF_ORDER_RECORD,
                            RTGRFF_S_PER_STEP, 0, (double *)slab, (double *)slab + plane, 1, stats));
    double *image = malloc(2 * (size_t)n_pix * n_pix * sizeof(double));
    CHECK(rtgrff_gather_image(ctx, (const double *)slab, 2, n_pix, n_pix, 0, image, 0));
    printf("# %s: %d x %d pixels, %d^3 cube, %.6g Hz; nominal %lld active %lld ray-steps\n", rtgrff_version(), n_pix, n_pix, n,
           freq, (long long)stats[0], (long long)stats[1]);
    for (size_t q = 0; q < 2 * (size_t)n_pix * n_pix; ++q) printf("%.17g\n", image[q]);

    CHECK(rtgrff_device_free(ctx, slab));
    CHECK(rtgrff_comm_destroy(ctx));
    CHECK(rtgrff_ctx_destroy(ctx));
    free(grid); free(omega); free(ne); free(te); free(bb); free(rows); free(xs); free(ys); free(zs); free(image);
    return 0;
}
