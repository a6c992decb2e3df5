// C-ABI of librtgrff_b200.so (declared in include/rtgrff.h).  Host-side staging + launches only;
// the kernels live in the .cuh files next to this one.
#include <stdarg.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "cube_builder.cuh"
#include "fused_map.cuh"
#include "grff.cuh"
#include "image_ops.cuh"
#include "los_sampler.cuh"
#include "ray_integrator.cuh"
#include "trace_kernel.cuh"

namespace rtgrff {

thread_local char g_err[512] = "";

inline int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap) return RTGRFF_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        p = nullptr;
        return fail(e == cudaErrorMemoryAllocation ? RTGRFF_ENOMEM : RTGRFF_ECUDA, "cudaMalloc(%zu) -> %s", bytes,
                    cudaGetErrorString(e));
    }
    cap = bytes;
    return RTGRFF_OK;
}

void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static int geom_from(const double g[12], int nx, int ny, int nz, GridGeom &o)
{
    if (nx < 2 || ny < 2 || nz < 2) return fail(RTGRFF_EINVAL, "cube needs >= 2 points per axis (%d,%d,%d)", nx, ny, nz);
    if ((int64_t)nx * ny * nz >= (int64_t)1 << 31) return fail(RTGRFF_EINVAL, "cube larger than 2^31 voxels");
    for (int a = 0; a < 3; ++a)
        if (!(g[4 * a + 1] > 0.0) || !isfinite(g[4 * a + 1])) return fail(RTGRFF_EINVAL, "axis %d has invalid spacing", a);
    o.nx = nx; o.ny = ny; o.nz = nz;
    o.x0 = g[0]; o.idx = 1.0 / g[1]; o.xl = g[2];
    o.y0 = g[4]; o.idy = 1.0 / g[5]; o.yl = g[6];
    o.z0 = g[8]; o.idz = 1.0 / g[9]; o.zl = g[10];
    return RTGRFF_OK;
}

static RayCube ray_cube_of(const rtgrff_ctx *c)
{
    RayCube r;
    const GridGeom &g = c->wgeom;
    r.c = c->wcube.as<float4>();
    r.pc = c->has_pcube ? c->pcube.as<float4>() : nullptr;
    r.nx = g.nx; r.ny = g.ny; r.nz = g.nz;
    r.sy = g.nz; r.sx = g.ny * g.nz;
    r.x0 = g.x0; r.y0 = g.y0; r.z0 = g.z0;
    r.xl = g.xl; r.yl = g.yl; r.zl = g.zl;
    r.idx = g.idx; r.idy = g.idy; r.idz = g.idz;
    return r;
}

static int use(rtgrff_ctx *c)
{
    if (!c) return fail(RTGRFF_EINVAL, "null context");
    RT_CUDA(cudaSetDevice(c->device));
    return RTGRFF_OK;
}

static int h2d(rtgrff_ctx *c, DevBuf &b, const void *src, size_t bytes)
{
    RT_TRY(b.reserve(bytes ? bytes : 1));
    if (bytes) RT_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return RTGRFF_OK;
}

static int d2h(rtgrff_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (bytes) RT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return RTGRFF_OK;
}

static int launched(rtgrff_ctx *c, const char *what)
{
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RTGRFF_ECUDA, "launch %s -> %s", what, cudaGetErrorString(e));
    return RTGRFF_OK;
}

// Cell-major polynomial cube for the FP32 stepper: 128 B per cell (8x the node cube).  Built when it
// fits in a third of the free device memory (RTGRFF_POLY_CUBE=0 disables it, =1 forces it); the
// stepper falls back to differencing the node cube on the fly without it.
static int build_poly_cube(rtgrff_ctx *c, int nx, int ny, int nz)
{
    c->has_pcube = false;
    const char *e = getenv("RTGRFF_POLY_CUBE");
    if (e && e[0] == '0') return RTGRFF_OK;
    const size_t bytes = (size_t)nx * ny * nz * 8 * sizeof(float4);
    if (bytes > c->pcube.cap && !(e && e[0] == '1')) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || bytes > free_b / 3) return RTGRFF_OK;
    }
    RT_TRY(c->pcube.reserve(bytes));
    build_poly_cube_kernel<<<blocks_for((int64_t)nx * ny * nz, 256), 256, 0, c->stream>>>(c->wcube.as<float4>(),
                                                                                         c->pcube.as<float4>(), nx, ny, nz);
    RT_TRY(launched(c, "build_poly_cube_kernel"));
    c->has_pcube = true;
    return RTGRFF_OK;
}


static int cs_every_step()
{
    // RTGRFF_CS_EVERY_STEP=1: trace the two cross-section rays at every step like the reference does,
    // although only the ratio of a recorded step is ever output (A/B switch; results are identical).
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RTGRFF_CS_EVERY_STEP");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

static int trace_variant()
{
    // RTGRFF_MODE selects the ray stepper: 0 (default) FP64 master state + FP32 cell-relative RHS with a
    // register cell cache; 1 FP64 state and RHS, FP32 trilinear arithmetic; 2 FP64 everything.
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RTGRFF_MODE");
        v = (e && e[0] >= '0' && e[0] <= '2') ? (e[0] - '0') : 0;
    }
    return v;
}

}  // namespace rtgrff

using namespace rtgrff;

extern "C" {

const char *rtgrff_version(void) { return "rtgrff_b200 0.1.0 sm_100a"; }
const char *rtgrff_last_error(void) { return g_err; }

int rtgrff_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(RTGRFF_ECUDA, "cudaGetDeviceCount -> %s", cudaGetErrorString(e));
    return n;
}

int rtgrff_ctx_create(int device, void *stream, rtgrff_ctx **out)
{
    if (!out) return fail(RTGRFF_EINVAL, "out is null");
    *out = nullptr;
    int n = 0;
    RT_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(RTGRFF_EINVAL, "device %d out of range (%d visible)", device, n);
    RT_CUDA(cudaSetDevice(device));
    rtgrff_ctx *c = new (std::nothrow) rtgrff_ctx();
    if (!c) return fail(RTGRFF_ENOMEM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        RT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    RT_CUDA(cudaEventCreate(&c->ev0));
    RT_CUDA(cudaEventCreate(&c->ev1));
    *out = c;
    return RTGRFF_OK;
}

int rtgrff_ctx_destroy(rtgrff_ctx *c)
{
    if (!c) return RTGRFF_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->wcube, &c->pcube, &c->fcube, &c->bcube, &c->rec_pos, &c->rec_s, &c->smp_ne, &c->smp_te, &c->smp_b,
                      &c->smp_ds, &c->smp_s, &c->smp_valid, &c->in0, &c->in1, &c->in2, &c->in3, &c->out0, &c->out1,
                      &c->out2, &c->out3, &c->out4, &c->out5, &c->stage, &c->counters};
    for (DevBuf *b : bufs) b->release();
    for (DevBuf &b : c->slot) b.release();
    c->slot_grids.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return RTGRFF_OK;
}

int rtgrff_ctx_synchronize(rtgrff_ctx *c)
{
    RT_TRY(use(c));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return RTGRFF_OK;
}

int64_t rtgrff_ctx_launch_count(const rtgrff_ctx *c) { return c ? c->launches : 0; }

double rtgrff_ctx_last_kernel_ms(rtgrff_ctx *c)
{
    if (!c || !c->ev_valid) return -1.0;
    float ms = -1.0f;
    if (cudaSetDevice(c->device) != cudaSuccess) return -1.0;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

int rtgrff_set_omega_cube(rtgrff_ctx *c, const double *omega_pe, int nx, int ny, int nz, const double geom[12],
                          int on_device)
{
    RT_TRY(use(c));
    if (!omega_pe || !geom) return fail(RTGRFF_EINVAL, "null argument");
    GridGeom g;
    RT_TRY(geom_from(geom, nx, ny, nz, g));
    const size_t nvox = (size_t)nx * ny * nz;
    const double *src = omega_pe;
    if (!on_device) {
        RT_TRY(h2d(c, c->stage, omega_pe, nvox * sizeof(double)));
        src = c->stage.as<double>();
    }
    RT_TRY(c->wcube.reserve(nvox * sizeof(float4)));
    build_ray_cube_kernel<<<blocks_for((int64_t)nvox, 256), 256, 0, c->stream>>>(src, c->wcube.as<float4>(), nx, ny, nz,
                                                                                 geom[3], geom[7], geom[11]);
    RT_TRY(launched(c, "build_ray_cube_kernel"));
    RT_TRY(build_poly_cube(c, nx, ny, nz));
    c->wgeom = g;
    c->has_wcube = true;
    RT_CUDA(cudaStreamSynchronize(c->stream));   // the host buffer may be reused by the caller
    return RTGRFF_OK;
}

int rtgrff_set_field_cubes(rtgrff_ctx *c, const float *ne, const float *te, const float *b, const float *bx,
                           const float *by, const float *bz, int nx, int ny, int nz, const double geom[12])
{
    RT_TRY(use(c));
    if (!ne || !te || !b || !geom) return fail(RTGRFF_EINVAL, "null argument");
    const bool bvec = bx && by && bz;
    if ((bx || by || bz) && !bvec) return fail(RTGRFF_EINVAL, "bx, by, bz must be given together");
    GridGeom g;
    RT_TRY(geom_from(geom, nx, ny, nz, g));
    const size_t nvox = (size_t)nx * ny * nz, fb = nvox * sizeof(float);
    RT_TRY(c->stage.reserve(3 * fb));
    float *s0 = c->stage.as<float>(), *s1 = s0 + nvox, *s2 = s1 + nvox;
    RT_TRY(c->fcube.reserve(nvox * sizeof(float4)));
    RT_CUDA(cudaMemcpyAsync(s0, ne, fb, cudaMemcpyHostToDevice, c->stream));
    RT_CUDA(cudaMemcpyAsync(s1, te, fb, cudaMemcpyHostToDevice, c->stream));
    RT_CUDA(cudaMemcpyAsync(s2, b, fb, cudaMemcpyHostToDevice, c->stream));
    interleave3_kernel<<<blocks_for((int64_t)nvox, 256), 256, 0, c->stream>>>(s0, s1, s2, c->fcube.as<float4>(), (int64_t)nvox);
    RT_TRY(launched(c, "interleave3_kernel"));
    if (bvec) {
        RT_TRY(c->bcube.reserve(nvox * sizeof(float4)));
        RT_CUDA(cudaMemcpyAsync(s0, bx, fb, cudaMemcpyHostToDevice, c->stream));
        RT_CUDA(cudaMemcpyAsync(s1, by, fb, cudaMemcpyHostToDevice, c->stream));
        RT_CUDA(cudaMemcpyAsync(s2, bz, fb, cudaMemcpyHostToDevice, c->stream));
        interleave3_kernel<<<blocks_for((int64_t)nvox, 256), 256, 0, c->stream>>>(s0, s1, s2, c->bcube.as<float4>(), (int64_t)nvox);
        RT_TRY(launched(c, "interleave3_kernel"));
    }
    c->fgeom = g;
    GridGeomF &f = c->fgeomf;
    f.nx = nx; f.ny = ny; f.nz = nz;
    f.x0 = (float)geom[0]; f.y0 = (float)geom[4]; f.z0 = (float)geom[8];
    f.idx = (float)(1.0 / geom[1]); f.idy = (float)(1.0 / geom[5]); f.idz = (float)(1.0 / geom[9]);
    f.fxl = (float)(nx - 1); f.fyl = (float)(ny - 1); f.fzl = (float)(nz - 1);
    c->has_fcube = true;
    c->has_bvec = bvec;
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return RTGRFF_OK;
}

int rtgrff_resample_spherical(rtgrff_ctx *c, int slot, const float *data, const double *phi, const double *lat,
                              const double *r, int np, int nt, int nr, const double *x_grid, const double *y_grid,
                              const double *z_grid, int nx, int ny, int nz, const double geom[12],
                              double phi0_offset_deg, double r_min, double scale, double fill, int fill_nonfinite,
                              double *out)
{
    RT_TRY(use(c));
    if (slot < 0 || slot > 4) return fail(RTGRFF_EINVAL, "slot %d out of range 0..4", slot);
    if (!data || !phi || !lat || !r || !geom || !x_grid || !y_grid || !z_grid) return fail(RTGRFF_EINVAL, "null argument");
    if (np < 2 || nt < 2 || nr < 2) return fail(RTGRFF_EINVAL, "spherical mesh needs >= 2 nodes per axis");
    GridGeom g;
    RT_TRY(geom_from(geom, nx, ny, nz, g));
    for (int i = 1; i < np; ++i) if (!(phi[i] > phi[i - 1])) return fail(RTGRFF_EINVAL, "phi nodes must ascend");
    for (int i = 1; i < nt; ++i) if (!(lat[i] > lat[i - 1])) return fail(RTGRFF_EINVAL, "latitude nodes must ascend");
    for (int i = 1; i < nr; ++i) if (!(r[i] > r[i - 1])) return fail(RTGRFF_EINVAL, "radius nodes must ascend");
    if (c->slot_nx && (c->slot_nx != nx || c->slot_ny != ny || c->slot_nz != nz))
        for (bool &b : c->slot_set) b = false;           // a new cube shape invalidates the other slots
    const size_t nvox = (size_t)nx * ny * nz, nd = (size_t)np * nt * nr;
    RT_TRY(h2d(c, c->in0, data, nd * sizeof(float)));
    RT_TRY(c->in1.reserve((size_t)(np + nt + nr) * sizeof(double)));
    double *ax = c->in1.as<double>();
    RT_CUDA(cudaMemcpyAsync(ax, phi, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RT_CUDA(cudaMemcpyAsync(ax + np, lat, nt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RT_CUDA(cudaMemcpyAsync(ax + np + nt, r, nr * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RT_TRY(c->slot[slot].reserve(nvox * sizeof(double)));
    RT_TRY(c->slot_grids.reserve((size_t)(nx + ny + nz) * sizeof(double)));
    double *gx = c->slot_grids.as<double>();
    RT_CUDA(cudaMemcpyAsync(gx, x_grid, nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RT_CUDA(cudaMemcpyAsync(gx + nx, y_grid, ny * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RT_CUDA(cudaMemcpyAsync(gx + nx + ny, z_grid, nz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ResampleArgs a;
    a.mesh.data = c->in0.as<float>();
    a.mesh.phi = ax; a.mesh.lat = ax + np; a.mesh.r = ax + np + nt;
    a.mesh.np = np; a.mesh.nt = nt; a.mesh.nr = nr;
    a.nx = nx; a.ny = ny; a.nz = nz;
    a.xg = gx; a.yg = gx + nx; a.zg = gx + nx + ny;
    a.phi0_offset_rad = phi0_offset_deg * M_PI / 180.0;
    a.r_min = r_min; a.scale = scale; a.fill = fill; a.fill_nonfinite = fill_nonfinite;
    a.out = c->slot[slot].as<double>();
    unsigned int blocks = blocks_for((int64_t)nvox, 256);
    const unsigned int cap = (unsigned int)c->sm_count * 32;
    if (blocks > cap) blocks = cap;
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
    resample_spherical_kernel<<<blocks, 256, 0, c->stream>>>(a);
    RT_TRY(launched(c, "resample_spherical_kernel"));
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    if (out) RT_TRY(d2h(c, out, c->slot[slot].p, nvox * sizeof(double)));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    c->slot_set[slot] = true;
    c->slot_nx = nx; c->slot_ny = ny; c->slot_nz = nz;
    memcpy(c->slot_geom, geom, sizeof(c->slot_geom));
    return RTGRFF_OK;
}

int rtgrff_sample_spherical_los(rtgrff_ctx *c, const float *data, const double *phi, const double *lat, const double *r,
                                int np, int nt, int nr, const double *x, const double *y, const double *zc, int nx,
                                int ny, int nz, double phi0_offset_deg, double r_min, double scale, double z_eps,
                                double *out)
{
    RT_TRY(use(c));
    if (!data || !phi || !lat || !r || !x || !y || !zc || !out) return fail(RTGRFF_EINVAL, "null argument");
    if (np < 2 || nt < 2 || nr < 2 || nx < 1 || ny < 1 || nz < 1) return fail(RTGRFF_EINVAL, "bad sizes");
    for (int i = 1; i < np; ++i) if (!(phi[i] > phi[i - 1])) return fail(RTGRFF_EINVAL, "phi nodes must ascend");
    for (int i = 1; i < nt; ++i) if (!(lat[i] > lat[i - 1])) return fail(RTGRFF_EINVAL, "latitude nodes must ascend");
    for (int i = 1; i < nr; ++i) if (!(r[i] > r[i - 1])) return fail(RTGRFF_EINVAL, "radius nodes must ascend");
    const size_t total = (size_t)nx * ny * nz, nd = (size_t)np * nt * nr;
    RT_TRY(h2d(c, c->in0, data, nd * sizeof(float)));
    RT_TRY(c->in1.reserve((size_t)(np + nt + nr + nx + ny + nz) * sizeof(double)));
    double *ax = c->in1.as<double>();
    const double *src[6] = {phi, lat, r, x, y, zc};
    const int len[6] = {np, nt, nr, nx, ny, nz};
    double *dst[6];
    size_t o = 0;
    for (int q = 0; q < 6; ++q) {
        dst[q] = ax + o;
        RT_CUDA(cudaMemcpyAsync(dst[q], src[q], len[q] * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        o += len[q];
    }
    RT_TRY(c->out0.reserve(total * sizeof(double)));
    LosArgs a;
    a.mesh.data = c->in0.as<float>();
    a.mesh.phi = dst[0]; a.mesh.lat = dst[1]; a.mesh.r = dst[2];
    a.mesh.np = np; a.mesh.nt = nt; a.mesh.nr = nr;
    a.x = dst[3]; a.y = dst[4]; a.zc = dst[5];
    a.nx = nx; a.ny = ny; a.nz = nz;
    a.phi0_offset_rad = phi0_offset_deg * M_PI / 180.0;
    a.r_min = r_min; a.scale = scale; a.z_eps = z_eps;
    a.out = c->out0.as<double>();
    unsigned int blocks = blocks_for((int64_t)total, 256);
    const unsigned int cap = (unsigned int)c->sm_count * 32;
    if (blocks > cap) blocks = cap;
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
    los_spherical_kernel<<<blocks, 256, 0, c->stream>>>(a);
    RT_TRY(launched(c, "los_spherical_kernel"));
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    RT_TRY(d2h(c, out, c->out0.p, total * sizeof(double)));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return RTGRFF_OK;
}

int rtgrff_compose_cubes(rtgrff_ctx *c, int want_bvec)
{
    RT_TRY(use(c));
    for (int s = 0; s < 5; ++s)
        if (!c->slot_set[s]) return fail(RTGRFF_ENOCUBE, "spherical slot %d (0 rho, 1 te, 2 br, 3 bt, 4 bp) is not set", s);
    const int nx = c->slot_nx, ny = c->slot_ny, nz = c->slot_nz;
    const double *geom = c->slot_geom;
    GridGeom g;
    RT_TRY(geom_from(geom, nx, ny, nz, g));
    const size_t nvox = (size_t)nx * ny * nz;
    RT_TRY(c->stage.reserve(nvox * sizeof(double)));
    RT_TRY(c->fcube.reserve(nvox * sizeof(float4)));
    if (want_bvec) RT_TRY(c->bcube.reserve(nvox * sizeof(float4)));
    RT_TRY(c->wcube.reserve(nvox * sizeof(float4)));
    ComposeArgs a;
    a.ne = c->slot[0].as<double>(); a.te = c->slot[1].as<double>();
    a.br = c->slot[2].as<double>(); a.bt = c->slot[3].as<double>(); a.bp = c->slot[4].as<double>();
    a.nx = nx; a.ny = ny; a.nz = nz;
    a.xg = c->slot_grids.as<double>(); a.yg = a.xg + nx; a.zg = a.yg + ny;
    a.omega_pe = c->stage.as<double>();
    a.fcube = c->fcube.as<float4>();
    a.bcube = want_bvec ? c->bcube.as<float4>() : nullptr;
    unsigned int blocks = blocks_for((int64_t)nvox, 256);
    compose_cubes_kernel<<<blocks, 256, 0, c->stream>>>(a);
    RT_TRY(launched(c, "compose_cubes_kernel"));
    build_ray_cube_kernel<<<blocks, 256, 0, c->stream>>>(c->stage.as<double>(), c->wcube.as<float4>(), nx, ny, nz,
                                                          geom[3], geom[7], geom[11]);
    RT_TRY(launched(c, "build_ray_cube_kernel"));
    RT_TRY(build_poly_cube(c, nx, ny, nz));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    c->wgeom = g; c->fgeom = g;
    GridGeomF &f = c->fgeomf;
    f.nx = nx; f.ny = ny; f.nz = nz;
    f.x0 = (float)geom[0]; f.y0 = (float)geom[4]; f.z0 = (float)geom[8];
    f.idx = (float)(1.0 / geom[1]); f.idy = (float)(1.0 / geom[5]); f.idz = (float)(1.0 / geom[9]);
    f.fxl = (float)(nx - 1); f.fyl = (float)(ny - 1); f.fzl = (float)(nz - 1);
    c->has_wcube = true; c->has_fcube = true; c->has_bvec = want_bvec != 0;
    return RTGRFF_OK;
}

int rtgrff_trace(rtgrff_ctx *c, int64_t n_rays, const double *x_start, const double *y_start, const double *z_start,
                 const double *kvec, double freq_hz, double dt, int64_t n_steps, int64_t record_stride, int trace_cs,
                 double perturb_ratio, int s_mode, double *r_record, double *s_record, int64_t *active_steps)
{
    RT_TRY(use(c));
    if (!c->has_wcube) return fail(RTGRFF_ENOCUBE, "rtgrff_set_omega_cube has not been called");
    if (n_rays < 0 || n_steps < 0 || record_stride < 1) return fail(RTGRFF_EINVAL, "bad n_rays/n_steps/record_stride");
    if (n_rays > 0 && (!x_start || !y_start || !z_start)) return fail(RTGRFF_EINVAL, "null start arrays");
    const int64_t n_rec = n_steps > 0 ? (n_steps + record_stride - 1) / record_stride : 0;
    c->rec_n = n_rec; c->rec_rays = n_rays; c->rec_has_s = trace_cs != 0;
    if (active_steps) *active_steps = 0;
    if (n_rays == 0 || n_rec == 0) return RTGRFF_OK;
    const size_t nb = (size_t)n_rays * sizeof(double);
    RT_TRY(h2d(c, c->in0, x_start, nb));
    RT_TRY(h2d(c, c->in1, y_start, nb));
    RT_TRY(h2d(c, c->in2, z_start, nb));
    if (kvec) RT_TRY(h2d(c, c->in3, kvec, 3 * nb));
    RT_TRY(c->rec_pos.reserve((size_t)n_rec * 3 * nb));
    if (trace_cs) RT_TRY(c->rec_s.reserve((size_t)n_rec * nb));
    RT_TRY(c->counters.reserve(64));
    RT_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));

    TraceArgs a;
    a.cube = ray_cube_of(c);
    a.n_rays = n_rays;
    a.x_start = c->in0.as<double>(); a.y_start = c->in1.as<double>(); a.z_start = c->in2.as<double>();
    a.kvec = kvec ? c->in3.as<double>() : nullptr;
    a.omega0 = 2.0 * M_PI * freq_hz;
    a.dt = dt; a.perturb_ratio = perturb_ratio;
    a.K = make_step_const(a.cube, dt, perturb_ratio);
    a.n_steps = n_steps; a.stride = record_stride; a.n_rec = n_rec;
    a.s_mode = s_mode;
    a.cs_every_step = cs_every_step();
    a.rec_pos = c->rec_pos.as<double>();
    a.rec_s = trace_cs ? c->rec_s.as<double>() : nullptr;
    a.active_steps = c->counters.as<unsigned long long>();
    const dim3 grid(blocks_for(n_rays, RT_BLOCK)), block(RT_BLOCK);
    int l64 = trace_variant();
    // the FP32 cell-relative stepper keeps every stage within a few cells of the cached cell
    if (l64 == MODE_FAST32 &&
        !(max_stage_offset_cells(dt, trace_cs ? perturb_ratio : 0.0, c->wgeom.idx, c->wgeom.idy, c->wgeom.idz) < kMaxStageOffsetCells))
        l64 = MODE_F64;
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
#define RT_TRACE_CASE(v, CS, M) \
    case v: trace_rays_kernel<CS, M><<<grid, block, 0, c->stream>>>(a); break;
    switch ((trace_cs ? 3 : 0) + l64) {
        RT_TRACE_CASE(0, false, 0) RT_TRACE_CASE(1, false, 1) RT_TRACE_CASE(2, false, 2)
        RT_TRACE_CASE(3, true, 0) RT_TRACE_CASE(4, true, 1) RT_TRACE_CASE(5, true, 2)
    }
#undef RT_TRACE_CASE
    RT_TRY(launched(c, "trace_rays_kernel"));
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    if (r_record) {
        RT_TRY(c->out0.reserve((size_t)n_rec * 3 * nb));
        records_to_aos_kernel<<<blocks_for(n_rec * n_rays * 3, 256), 256, 0, c->stream>>>(
            c->rec_pos.as<double>(), c->out0.as<double>(), n_rec, n_rays);
        RT_TRY(launched(c, "records_to_aos_kernel"));
        RT_TRY(d2h(c, r_record, c->out0.p, (size_t)n_rec * 3 * nb));
    }
    if (s_record && trace_cs) RT_TRY(d2h(c, s_record, c->rec_s.p, (size_t)n_rec * nb));
    unsigned long long act = 0;
    RT_TRY(d2h(c, &act, c->counters.p, sizeof(act)));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    if (active_steps) *active_steps = (int64_t)act;
    return RTGRFF_OK;
}

static int run_sampler(rtgrff_ctx *c, SampleArgs &a, double r_sun_cm, double fill_ne, double fill_te, double fill_b)
{
    const size_t n = (size_t)a.n_rec * a.n_rays;
    RT_TRY(c->smp_ne.reserve(n * 4)); RT_TRY(c->smp_te.reserve(n * 4)); RT_TRY(c->smp_b.reserve(n * 4));
    RT_TRY(c->smp_ds.reserve(n * 4)); RT_TRY(c->smp_s.reserve(n * 4)); RT_TRY(c->smp_valid.reserve(n));
    a.fcube = c->fcube.as<float4>();
    a.g = c->fgeomf;
    a.r_sun_cm = (float)r_sun_cm;   // numpy 2: float32 array * python float stays float32 (gpu_raytrace.py:482-484)
    a.fill_ne = (float)fill_ne; a.fill_te = (float)fill_te; a.fill_b = (float)fill_b;
    a.ne = c->smp_ne.as<float>(); a.te = c->smp_te.as<float>(); a.b = c->smp_b.as<float>();
    a.ds = c->smp_ds.as<float>(); a.s_out = c->smp_s.as<float>(); a.valid = c->smp_valid.as<uint8_t>();
    unsigned int blocks = blocks_for((int64_t)n, 256);
    const unsigned int cap = (unsigned int)c->sm_count * 64;
    if (blocks > cap) blocks = cap;
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
    sample_paths_kernel<<<blocks, 256, 0, c->stream>>>(a);
    RT_TRY(launched(c, "sample_paths_kernel"));
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    c->smp_n = a.n_rec; c->smp_rays = a.n_rays;
    return RTGRFF_OK;
}

static int sampler_out(rtgrff_ctx *c, size_t n, float *ne, float *te, float *b, float *ds, uint8_t *valid, float *s)
{
    if (ne) RT_TRY(d2h(c, ne, c->smp_ne.p, n * 4));
    if (te) RT_TRY(d2h(c, te, c->smp_te.p, n * 4));
    if (b) RT_TRY(d2h(c, b, c->smp_b.p, n * 4));
    if (ds) RT_TRY(d2h(c, ds, c->smp_ds.p, n * 4));
    if (valid) RT_TRY(d2h(c, valid, c->smp_valid.p, n));
    if (s) RT_TRY(d2h(c, s, c->smp_s.p, n * 4));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return RTGRFF_OK;
}

int rtgrff_sample(rtgrff_ctx *c, int64_t n_rec, int64_t n_rays, const float *pos, const float *s,
                  const float *ray_start, double r_sun_cm, double fill_ne, double fill_te, double fill_b, float *ne,
                  float *te, float *b, float *ds, uint8_t *valid)
{
    RT_TRY(use(c));
    if (!c->has_fcube) return fail(RTGRFF_ENOCUBE, "rtgrff_set_field_cubes has not been called");
    if (n_rec < 0 || n_rays < 0) return fail(RTGRFF_EINVAL, "negative sizes");
    const size_t n = (size_t)n_rec * n_rays;
    if (n == 0) return RTGRFF_OK;
    if (!pos || !s || !ray_start) return fail(RTGRFF_EINVAL, "null input");
    RT_TRY(h2d(c, c->in0, pos, n * 3 * sizeof(float)));
    RT_TRY(h2d(c, c->in1, s, n * sizeof(float)));
    RT_TRY(h2d(c, c->in2, ray_start, (size_t)n_rays * 3 * sizeof(float)));
    SampleArgs a{};
    a.n_rec = n_rec; a.n_rays = n_rays;
    a.pos_aos = c->in0.as<float>(); a.s32 = c->in1.as<float>(); a.ray_start = c->in2.as<float>();
    RT_TRY(run_sampler(c, a, r_sun_cm, fill_ne, fill_te, fill_b));
    return sampler_out(c, n, ne, te, b, ds, valid, nullptr);
}

int rtgrff_sample_traced(rtgrff_ctx *c, const float *ray_start, double r_sun_cm, double fill_ne, double fill_te,
                         double fill_b, float *ne, float *te, float *b, float *ds, uint8_t *valid, float *s)
{
    RT_TRY(use(c));
    if (!c->has_fcube) return fail(RTGRFF_ENOCUBE, "rtgrff_set_field_cubes has not been called");
    if (c->rec_n <= 0 || c->rec_rays <= 0) return fail(RTGRFF_EINVAL, "no traced records on the device");
    if (!ray_start) return fail(RTGRFF_EINVAL, "null ray_start");
    const size_t n = (size_t)c->rec_n * c->rec_rays;
    RT_TRY(h2d(c, c->in2, ray_start, (size_t)c->rec_rays * 3 * sizeof(float)));
    SampleArgs a{};
    a.n_rec = c->rec_n; a.n_rays = c->rec_rays;
    a.pos_soa = c->rec_pos.as<double>();
    a.s64 = c->rec_has_s ? c->rec_s.as<double>() : nullptr;
    a.ray_start = c->in2.as<float>();
    RT_TRY(run_sampler(c, a, r_sun_cm, fill_ne, fill_te, fill_b));
    return sampler_out(c, n, ne, te, b, ds, valid, s);
}

static int run_slice(rtgrff_ctx *c, int npix, int nz, int nf, const double *rparms, const double *parms, double *rl,
                     int32_t *status)
{
    const size_t pb = (size_t)15 * nz * npix * sizeof(double), rb = (size_t)3 * npix * sizeof(double);
    const size_t ob = (size_t)7 * nf * npix * sizeof(double);
    RT_TRY(h2d(c, c->in0, parms, pb));
    RT_TRY(h2d(c, c->in1, rparms, rb));
    RT_TRY(c->out0.reserve(ob));
    RT_TRY(c->out1.reserve((size_t)npix * sizeof(int32_t)));
    RT_CUDA(cudaMemsetAsync(c->out1.p, 0xff, (size_t)npix * sizeof(int32_t), c->stream));
    SliceArgs a;
    a.parms = c->in0.as<double>(); a.rparms = c->in1.as<double>();
    a.rl = c->out0.as<double>(); a.status = c->out1.as<int32_t>();
    a.npix = npix; a.nz = nz; a.nf = nf;
    const int64_t warps = (int64_t)npix * nf;
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
    grff_slice_kernel<<<blocks_for(warps * 32, 128), 128, 0, c->stream>>>(a);
    RT_TRY(launched(c, "grff_slice_kernel"));
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    RT_TRY(d2h(c, rl, c->out0.p, ob));
    if (status) RT_TRY(d2h(c, status, c->out1.p, (size_t)npix * sizeof(int32_t)));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return RTGRFF_OK;
}

int rtgrff_get_mw_slice(rtgrff_ctx *c, const int32_t *Lparms_M, const double *Rparms_M, const double *Parms_M,
                        const double *T_arr, const double *DEM_arr, const double *DDM_arr, double *RL_M, int32_t *status)
{
    (void)T_arr; (void)DEM_arr; (void)DDM_arr;
    RT_TRY(use(c));
    if (!Lparms_M || !Rparms_M || !Parms_M || !RL_M) return fail(RTGRFF_EINVAL, "null argument");
    const int npix = Lparms_M[0], nz = Lparms_M[1], nf = Lparms_M[2];
    if (npix < 0 || nz < 0 || nf <= 0) return fail(RTGRFF_EINVAL, "bad Lparms_M {%d,%d,%d}", npix, nz, nf);
    if (npix == 0) return RTGRFF_OK;
    return run_slice(c, npix, nz, nf, Rparms_M, Parms_M, RL_M, status);
}

static std::mutex g_default_mu;
static rtgrff_ctx *g_default_ctx = nullptr;

int PyGET_MW(const int32_t *Lparms, const double *Rparms, const double *Parms, const double *T_arr,
             const double *DEM_arr, const double *DDM_arr, double *RL)
{
    (void)T_arr; (void)DEM_arr; (void)DDM_arr;
    if (!Lparms || !Rparms || !Parms || !RL) return 1;
    const int nz = Lparms[0], nf = Lparms[1];
    if (nz < 0 || nf <= 0) return 1;
    if (Lparms[2] > 0) return 2;
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_ctx && rtgrff_ctx_create(0, nullptr, &g_default_ctx) != RTGRFF_OK) return 3;
    if (use(g_default_ctx) != RTGRFF_OK) return 3;
    return run_slice(g_default_ctx, 1, nz, nf, Rparms, Parms, RL, nullptr) == RTGRFF_OK ? 0 : 3;
}

int rtgrff_emission_traced(rtgrff_ctx *c, double pixel_area_cm2, double freq0_hz, int n_freq, double freq_log_step,
                           int em_flag, int s_max, int s_input_on, double *tb, double *vi)
{
    RT_TRY(use(c));
    if (c->smp_n <= 0 || c->smp_rays <= 0) return fail(RTGRFF_EINVAL, "no samples on the device (call rtgrff_sample_traced)");
    if (n_freq <= 0 || !tb || !vi) return fail(RTGRFF_EINVAL, "bad n_freq / null output");
    const size_t n = (size_t)c->smp_rays * n_freq;
    RT_TRY(c->out0.reserve(n * sizeof(double)));
    RT_TRY(c->out1.reserve(n * sizeof(double)));
    EmissionArgs a;
    a.ne = c->smp_ne.as<float>(); a.te = c->smp_te.as<float>(); a.b = c->smp_b.as<float>(); a.ds = c->smp_ds.as<float>();
    a.valid = c->smp_valid.as<uint8_t>();
    a.s = c->smp_s.as<float>(); a.s_input = s_input_on ? 1 : 0;
    a.n_rec = c->smp_n; a.n_rays = c->smp_rays;
    a.area = pixel_area_cm2; a.freq0 = freq0_hz; a.log_step = freq_log_step;
    a.n_freq = n_freq; a.em_flag = em_flag; a.s_max = s_max;
    a.tb = c->out0.as<double>(); a.vi = c->out1.as<double>();
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
    emission_rays_kernel<<<blocks_for((int64_t)n, 128), 128, 0, c->stream>>>(a);
    RT_TRY(launched(c, "emission_rays_kernel"));
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    RT_TRY(d2h(c, tb, c->out0.p, n * sizeof(double)));
    RT_TRY(d2h(c, vi, c->out1.p, n * sizeof(double)));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return RTGRFF_OK;
}

int rtgrff_render_map(rtgrff_ctx *c, int64_t n_rays, const double *x_start, const double *y_start,
                      const double *z_start, const double *kvec, const int32_t *ray_order, int n_freq,
                      const rtgrff_freq_params *freqs,
                      int trace_cs, double perturb_ratio, double pixel_area_cm2, double r_sun_cm, int em_flag, int s_max,
                      int use_bvec, int voxel_order, int s_mode, int s_input_on, double *tb, double *vi, int out_on_device,
                      int64_t *stats)
{
    RT_TRY(use(c));
    if (!c->has_wcube) return fail(RTGRFF_ENOCUBE, "rtgrff_set_omega_cube has not been called");
    if (!c->has_fcube) return fail(RTGRFF_ENOCUBE, "rtgrff_set_field_cubes has not been called");
    if (use_bvec && !c->has_bvec) return fail(RTGRFF_ENOCUBE, "use_bvec needs bx,by,bz in rtgrff_set_field_cubes");
    if (n_rays < 0 || n_freq <= 0 || n_freq > 65535 || !freqs || !tb || !vi) return fail(RTGRFF_EINVAL, "bad arguments");
    if (voxel_order != RTGRFF_ORDER_RECORD && voxel_order != RTGRFF_ORDER_REVERSED) return fail(RTGRFF_EINVAL, "bad voxel_order");
    if (s_mode != RTGRFF_S_PER_STEP && s_mode != RTGRFF_S_CUMULATIVE) return fail(RTGRFF_EINVAL, "bad s_mode");
    if (s_input_on && !trace_cs) return fail(RTGRFF_EINVAL, "s_input_on needs the cross-sections traced (trace_cs)");
    if (n_rays > 0 && (!x_start || !y_start || !z_start)) return fail(RTGRFF_EINVAL, "null start arrays");
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (n_rays == 0) return RTGRFF_OK;
    const RayCube rc = ray_cube_of(c);
    std::vector<FreqDev> fd(n_freq);
    int64_t nominal = 0;
    for (int f = 0; f < n_freq; ++f) {
        if (freqs[f].n_steps < 0 || freqs[f].record_stride < 1 || !(freqs[f].freq_hz > 0.0) ||
            freqs[f].n_steps > 2147483000 || freqs[f].record_stride > 2147483000)
            return fail(RTGRFF_EINVAL, "bad per-frequency parameters at index %d", f);
        fd[f].nu = freqs[f].freq_hz;
        fd[f].omega0 = 2.0 * M_PI * freqs[f].freq_hz;
        fd[f].dt = freqs[f].dt;
        fd[f].n_steps = freqs[f].n_steps;
        fd[f].stride = freqs[f].record_stride;
        fd[f].out_row = f;
        fd[f].K = make_step_const(rc, freqs[f].dt, perturb_ratio);
        fd[f].fq = make_freq(freqs[f].freq_hz);
        nominal += freqs[f].n_steps * n_rays;
    }
    const size_t nb = (size_t)n_rays * sizeof(double);
    RT_TRY(h2d(c, c->in0, x_start, nb));
    RT_TRY(h2d(c, c->in1, y_start, nb));
    RT_TRY(h2d(c, c->in2, z_start, nb));
    if (kvec) RT_TRY(h2d(c, c->in3, kvec, 3 * nb));
    if (ray_order) {
        if (n_rays >= ((int64_t)1 << 31)) return fail(RTGRFF_EINVAL, "ray_order needs n_rays < 2^31");
        RT_TRY(h2d(c, c->out2, ray_order, (size_t)n_rays * sizeof(int32_t)));
    }
    RT_TRY(c->counters.reserve(64));
    RT_CUDA(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    double *dtb = tb, *dvi = vi;
    if (!out_on_device) {
        RT_TRY(c->out0.reserve((size_t)n_freq * nb));
        RT_TRY(c->out1.reserve((size_t)n_freq * nb));
        dtb = c->out0.as<double>(); dvi = c->out1.as<double>();
    }
    MapArgs a;
    a.cube = rc;
    a.fcube = c->fcube.as<float4>();
    a.bcube = c->has_bvec ? c->bcube.as<float4>() : nullptr;
    a.fg = c->fgeomf;
    a.n_rays = n_rays;
    a.x_start = c->in0.as<double>(); a.y_start = c->in1.as<double>(); a.z_start = c->in2.as<double>();
    a.kvec = kvec ? c->in3.as<double>() : nullptr;
    a.ray_order = ray_order ? c->out2.as<int>() : nullptr;
    a.perturb_ratio = perturb_ratio; a.area = pixel_area_cm2;
    a.r_sun_cm = (float)r_sun_cm; a.fill_ne = 0.0f; a.fill_te = 1e4f; a.fill_b = 0.0f;
    a.em_flag = em_flag; a.s_max = s_max; a.use_bvec = use_bvec; a.order = voxel_order;
    a.cs_every_step = cs_every_step();
    a.s_mode = s_mode; a.s_input = s_input_on ? 1 : 0;
    a.tb = dtb; a.vi = dvi;
    a.active_steps = c->counters.as<unsigned long long>();
    const dim3 block(RT_BLOCK);
    const bool gr = !(em_flag & 1);
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
    const int variant = (trace_cs ? 8 : 0) | (voxel_order == RTGRFF_ORDER_REVERSED ? 4 : 0) | (use_bvec ? 2 : 0) | (gr ? 1 : 0);
#define RT_MAP_CASE(v, CS, ORD, BV, GR)                                                                  \
    case v:                                                                                              \
        if (mode == MODE_FAST32) render_map_kernel<CS, ORD, BV, GR, MODE_FAST32><<<grid, block, 0, c->stream>>>(a); \
        else render_map_kernel<CS, ORD, BV, GR, MODE_F64><<<grid, block, 0, c->stream>>>(a);             \
        break;
    int mode = trace_variant() == 0 ? MODE_FAST32 : MODE_F64;
    for (int f = 0; f < n_freq; ++f)
        if (!(max_stage_offset_cells(freqs[f].dt, trace_cs ? perturb_ratio : 0.0, c->wgeom.idx, c->wgeom.idy,
                                     c->wgeom.idz) < kMaxStageOffsetCells))
            mode = MODE_F64;
    // Longest frequencies first: blocks are dispatched in grid order, and a launch that ends on its
    // most expensive blocks (15 000 steps with a record at every step) idles most of the chip in its tail.
    std::stable_sort(fd.begin(), fd.end(), [&](const FreqDev &x, const FreqDev &y) {
        auto cost = [&](const FreqDev &q) {
            return (double)q.n_steps * (1.0 + ((trace_cs ? 2.0 : 0.0) + 1.0) / (double)q.stride);
        };
        return cost(x) > cost(y);
    });
    // the per-frequency constants travel in the kernel parameters: kMaxFreqPerLaunch frequencies per launch
    for (int f0 = 0; f0 < n_freq; f0 += kMaxFreqPerLaunch) {
    a.n_freq = n_freq - f0 < kMaxFreqPerLaunch ? n_freq - f0 : kMaxFreqPerLaunch;
    for (int f = 0; f < a.n_freq; ++f) a.freqs[f] = fd[f0 + f];
    const dim3 grid(blocks_for(n_rays, RT_BLOCK), (unsigned int)a.n_freq);
    switch (variant) {
        RT_MAP_CASE(0, false, 0, false, false) RT_MAP_CASE(1, false, 0, false, true)
        RT_MAP_CASE(2, false, 0, true, false) RT_MAP_CASE(3, false, 0, true, true)
        RT_MAP_CASE(4, false, 1, false, false) RT_MAP_CASE(5, false, 1, false, true)
        RT_MAP_CASE(6, false, 1, true, false) RT_MAP_CASE(7, false, 1, true, true)
        RT_MAP_CASE(8, true, 0, false, false) RT_MAP_CASE(9, true, 0, false, true)
        RT_MAP_CASE(10, true, 0, true, false) RT_MAP_CASE(11, true, 0, true, true)
        RT_MAP_CASE(12, true, 1, false, false) RT_MAP_CASE(13, true, 1, false, true)
        RT_MAP_CASE(14, true, 1, true, false) RT_MAP_CASE(15, true, 1, true, true)
    }
    RT_TRY(launched(c, "render_map_kernel"));
    }
#undef RT_MAP_CASE
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    if (!out_on_device) {
        RT_TRY(d2h(c, tb, dtb, (size_t)n_freq * nb));
        RT_TRY(d2h(c, vi, dvi, (size_t)n_freq * nb));
    }
    unsigned long long act[3] = {0, 0, 0};
    RT_TRY(d2h(c, act, c->counters.p, sizeof(act)));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    if (stats) { stats[0] = nominal; stats[1] = (int64_t)act[0]; stats[2] = (int64_t)act[1]; stats[3] = (int64_t)act[2]; }
    return RTGRFF_OK;
}

int rtgrff_gaussian_beam(rtgrff_ctx *c, const double *img, int ny, int nx, int n_planes, double sigma_pix,
                         double truncate, double *out)
{
    RT_TRY(use(c));
    if (!img || !out || ny < 1 || nx < 1 || n_planes < 1) return fail(RTGRFF_EINVAL, "bad arguments");
    if (!(sigma_pix >= 0.0) || !isfinite(sigma_pix) || !(truncate > 0.0)) return fail(RTGRFF_EINVAL, "bad sigma/truncate");
    const size_t n = (size_t)n_planes * ny * nx, nb = n * sizeof(double);
    // scipy.ndimage._gaussian_kernel1d: radius = int(truncate*sigma + 0.5), exp(-x^2/(2 sigma^2)) / sum
    const int radius = (int)(truncate * sigma_pix + 0.5);
    std::vector<double> w(radius + 1);
    double sum = 0.0;
    for (int d = -radius; d <= radius; ++d) sum += sigma_pix > 0.0 ? exp(-0.5 / (sigma_pix * sigma_pix) * (double)d * (double)d) : (d == 0 ? 1.0 : 0.0);
    for (int d = 0; d <= radius; ++d)
        w[d] = (sigma_pix > 0.0 ? exp(-0.5 / (sigma_pix * sigma_pix) * (double)d * (double)d) : (d == 0 ? 1.0 : 0.0)) / sum;
    RT_TRY(h2d(c, c->in0, img, nb));
    RT_TRY(h2d(c, c->in1, w.data(), w.size() * sizeof(double)));
    RT_TRY(c->out0.reserve(nb));
    RT_TRY(c->out1.reserve(nb));
    const unsigned int blocks = (unsigned int)std::min<int64_t>(blocks_for((int64_t)n, 256), (int64_t)c->sm_count * 32);
    RT_CUDA(cudaEventRecord(c->ev0, c->stream));
    // gaussian_filter runs the 1-D filter axis by axis: rows direction (axis 0) first, then axis 1
    gaussian_pass_kernel<<<blocks, 256, 0, c->stream>>>(c->in0.as<double>(), c->out0.as<double>(), c->in1.as<double>(),
                                                        radius, ny, nx, n_planes, 0);
    RT_TRY(launched(c, "gaussian_pass_kernel"));
    gaussian_pass_kernel<<<blocks, 256, 0, c->stream>>>(c->out0.as<double>(), c->out1.as<double>(), c->in1.as<double>(),
                                                        radius, ny, nx, n_planes, 1);
    RT_TRY(launched(c, "gaussian_pass_kernel"));
    RT_CUDA(cudaEventRecord(c->ev1, c->stream));
    c->ev_valid = true;
    RT_TRY(d2h(c, out, c->out1.p, nb));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    return RTGRFF_OK;
}

int rtgrff_patch_nan(rtgrff_ctx *c, double *img, int ny, int nx, int n_planes, int max_passes, int64_t *n_patched)
{
    RT_TRY(use(c));
    if (!img || ny < 1 || nx < 1 || n_planes < 1 || max_passes < 0) return fail(RTGRFF_EINVAL, "bad arguments");
    if (n_patched) *n_patched = 0;
    const size_t n = (size_t)n_planes * ny * nx, nb = n * sizeof(double);
    RT_TRY(h2d(c, c->in0, img, nb));
    RT_TRY(c->out0.reserve(nb));
    RT_TRY(c->out1.reserve(nb));
    RT_TRY(c->out2.reserve(n));
    RT_TRY(c->counters.reserve(64 + (size_t)n_planes * sizeof(int)));
    int *d_any = c->counters.as<int>();
    int *d_fixed = reinterpret_cast<int *>(c->counters.as<char>() + 64);
    const unsigned int blocks = (unsigned int)std::min<int64_t>(blocks_for((int64_t)n, 256), (int64_t)c->sm_count * 32);
    std::vector<int> fixed(n_planes);
    int64_t total = 0;
    // _patch_nan_2d (util.py:45-76): up to max_passes sweeps; stop when nothing is left or nothing could be fixed
    for (int pass = 0; pass < max_passes; ++pass) {
        RT_CUDA(cudaMemsetAsync(c->counters.p, 0, 64 + (size_t)n_planes * sizeof(int), c->stream));
        mark_nonfinite_kernel<<<blocks, 256, 0, c->stream>>>(c->in0.as<double>(), c->out2.as<unsigned char>(), (int64_t)n, d_any);
        RT_TRY(launched(c, "mark_nonfinite_kernel"));
        int any = 0;
        RT_TRY(d2h(c, &any, d_any, sizeof(int)));
        RT_CUDA(cudaStreamSynchronize(c->stream));
        if (!any) break;
        next_finite_kernel<<<blocks_for((int64_t)n_planes * (ny + nx), 128), 128, 0, c->stream>>>(
            c->in0.as<double>(), c->out0.as<double>(), c->out1.as<double>(), ny, nx, n_planes);
        RT_TRY(launched(c, "next_finite_kernel"));
        patch_nan_pass_kernel<<<n_planes, 1024, 0, c->stream>>>(c->in0.as<double>(), c->out2.as<unsigned char>(),
                                                                c->out0.as<double>(), c->out1.as<double>(), ny, nx, d_fixed);
        RT_TRY(launched(c, "patch_nan_pass_kernel"));
        RT_TRY(d2h(c, fixed.data(), d_fixed, (size_t)n_planes * sizeof(int)));
        RT_CUDA(cudaStreamSynchronize(c->stream));
        int64_t f = 0;
        for (int v : fixed) f += v;
        total += f;
        if (f == 0) break;
    }
    RT_TRY(d2h(c, img, c->in0.p, nb));
    RT_CUDA(cudaStreamSynchronize(c->stream));
    if (n_patched) *n_patched = total;
    return RTGRFF_OK;
}

}  // extern "C"
