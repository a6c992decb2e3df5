// Ray stepper v2 for sm_100a: FP64 master state, FP32 cell-relative RHS, register cell cache,
// packed FP32x2 (FFMA2/FMUL2) trilinear arithmetic.
//
// Same mathematics as ray_integrator.cuh (build_rays.py:158-239) reorganised around what the
// first ncu capture showed (profiles/r1a_trace_rays_kernel_*.txt): the FP64 stepper issues ~2500
// instructions per ray-step at 49 % issue utilisation, with the L1 data pipe at 65 % (8 x LDG.128
// per RHS, 9.3 wavefronts each), the FP64 pipe at 29 % and the XU pipe (f64<->f32/int
// conversions, MUFU) at 29 %.
//
//  * The ray's position and wave vector stay FP64 (the "master" state): every step adds
//    (double)sum * (dt/6 * C_R), so the accumulation over thousands of steps is exact to 1e-16
//    and no rounded constant multiplies the accumulated path.
//  * Within a step everything is FP32 and CELL-RELATIVE: the master position is split once per
//    step into (base cell, fraction in [0,1)); the four RK4 stages and the two cross-section
//    rays are small float offsets from that cell (the host falls back to the FP64 stepper when
//    a step could span a whole cell).  Cell fractions carry 6e-8 of a cell (~1e-9 R_sun), the
//    corners are FP32 anyway, omega = sqrt(w^2+k^2) and the derivatives need 1e-7 relative:
//    increments are ~1e-3 R_sun, so the per-step error is ~1e-10 R_sun and random.
//  * The 8 corners of the current cell (32 floats) live in registers; a stage whose cell differs
//    from the cached one re-fetches with predicated LDG.128 (only the lanes that moved generate
//    L1 wavefronts).  A ray spends ~10 steps x 12 RHS evaluations in one cell.
//  * The four channels {omega_pe, d/dx, d/dy, d/dz} of a corner are two aligned float2 pairs, so
//    the 28 scalar lerps of a trilinear gather become 14 packed lerps = FMUL2 + FFMA2
//    (Blackwell's packed FP32x2 pipe): half the issue slots.
//  * Away from the cube faces no stage can leave the cube, so the per-stage bounds test is
//    hoisted to one test of the base cell per step (EDGE path only near the faces).
#pragma once

#include "ray_integrator.cuh"

namespace rtgrff {

struct Cell {
    int off;  // element offset of corner (i,j,k); -1 = empty
    float4 c000, c001, c010, c011, c100, c101, c110, c111;
};

struct Deriv32 {
    float vx, vy, vz;   // k / omega            (dr/dt = C_R * v)
    float gx, gy, gz;   // (omega_pe/omega) * grad omega_pe   (dk/dt = -C_R * g)
};

__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float2 u, float2 t)
{
    return __ffma2_rn(b, t, __fmul2_rn(a, u));   // a*(1-t) + b*t on both halves
}

// One RHS evaluation at cell-relative position p (cell units, -1 <= p < 2) of base cell (bi,bj,bk).
template <bool EDGE>
__device__ __forceinline__ Deriv32 rhs32(const RayCube &C, int bi, int bj, int bk, int base_off, Cell &cache, float px,
                                         float py, float pz, float kx, float ky, float kz)
{
    Deriv32 d;
    d.vx = d.vy = d.vz = d.gx = d.gy = d.gz = 0.0f;
    const bool hx = px >= 1.0f, lx = px < 0.0f, hy = py >= 1.0f, ly = py < 0.0f, hz = pz >= 1.0f, lz = pz < 0.0f;
    float tx = px, ty = py, tz = pz;
    tx = hx ? tx - 1.0f : tx; tx = lx ? tx + 1.0f : tx;
    ty = hy ? ty - 1.0f : ty; ty = ly ? ty + 1.0f : ty;
    tz = hz ? tz - 1.0f : tz; tz = lz ? tz + 1.0f : tz;
    int off;
    if (EDGE) {
        // scipy bounds per stage: g[0] <= x <= g[n-1]; the last node belongs to cell n-2 with t = 1
        int ci = bi + (int)hx - (int)lx, cj = bj + (int)hy - (int)ly, ck = bk + (int)hz - (int)lz;
        if (ci == C.nx - 1 && tx == 0.0f) { ci = C.nx - 2; tx = 1.0f; }
        if (cj == C.ny - 1 && ty == 0.0f) { cj = C.ny - 2; ty = 1.0f; }
        if (ck == C.nz - 1 && tz == 0.0f) { ck = C.nz - 2; tz = 1.0f; }
        if ((unsigned)ci > (unsigned)(C.nx - 2) || (unsigned)cj > (unsigned)(C.ny - 2) ||
            (unsigned)ck > (unsigned)(C.nz - 2))
            return d;
        off = (ci * C.ny + cj) * C.nz + ck;
    } else {
        off = base_off;
        off = hx ? off + C.sx : off; off = lx ? off - C.sx : off;
        off = hy ? off + C.sy : off; off = ly ? off - C.sy : off;
        off = hz ? off + 1 : off;    off = lz ? off - 1 : off;
    }
    if (off != cache.off) {
        const float4 *p = C.c + off;
        cache.c000 = __ldg(p); cache.c001 = __ldg(p + 1);
        cache.c010 = __ldg(p + C.sy); cache.c011 = __ldg(p + C.sy + 1);
        cache.c100 = __ldg(p + C.sx); cache.c101 = __ldg(p + C.sx + 1);
        cache.c110 = __ldg(p + C.sx + C.sy); cache.c111 = __ldg(p + C.sx + C.sy + 1);
        cache.off = off;
    }
    const float2 tz2 = make_float2(tz, tz), uz2 = make_float2(1.0f - tz, 1.0f - tz);
    const float2 ty2 = make_float2(ty, ty), uy2 = make_float2(1.0f - ty, 1.0f - ty);
    const float2 tx2 = make_float2(tx, tx), ux2 = make_float2(1.0f - tx, 1.0f - tx);
#define RT_LO(c) make_float2((c).x, (c).y)
#define RT_HI(c) make_float2((c).z, (c).w)
#define RT_TRI2(H)                                                                                          \
    lerp2(lerp2(lerp2(H(cache.c000), H(cache.c001), uz2, tz2), lerp2(H(cache.c010), H(cache.c011), uz2, tz2), uy2, ty2), \
          lerp2(lerp2(H(cache.c100), H(cache.c101), uz2, tz2), lerp2(H(cache.c110), H(cache.c111), uz2, tz2), uy2, ty2), ux2, tx2)
    const float2 wg = RT_TRI2(RT_LO);   // {omega_pe, d/dx}
    const float2 gg = RT_TRI2(RT_HI);   // {d/dy, d/dz}
#undef RT_TRI2
#undef RT_LO
#undef RT_HI
    const float w = wg.x;
    const float om2 = fmaf(w, w, fmaf(kx, kx, fmaf(ky, ky, kz * kz)));
    // valid = isfinite(omega_pe) & isfinite(omega) & (omega > 0)   (build_rays.py:169); a non-finite
    // omega_pe or k makes omega^2 non-finite, so one range test on omega^2 covers all three
    if (!(om2 > 0.0f) || !(om2 < INFINITY)) return d;
    float inv_om;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv_om) : "f"(om2));   // MUFU.RSQ, 2 ulp; omega^2 ~ 1e17
    const float a = w * inv_om;
    d.vx = kx * inv_om; d.vy = ky * inv_om; d.vz = kz * inv_om;
    d.gx = a * wg.y; d.gy = a * gg.x; d.gz = a * gg.y;
    return d;
}

// Per-launch float constants derived from dt and the grid (uniform across the grid).
struct StepConst {
    float hx, hy, hz;     // 0.5*dt*C_R/dx : half-step position offset per unit v, in cells
    float hk;             // 0.5*dt*C_R    : half-step k offset per unit g
    float c6r;            // dt/6*C_R      : step displacement per unit sum(v), R_sun
    float ix, iy, iz;     // 1/dx          : R_sun -> cells
    double c6;            // dt/6*C_R in double: master-state increments
    float perturb;
};

__device__ __forceinline__ StepConst make_step_const(const RayCube &C, double dt, double perturb_ratio)
{
    StepConst k;
    const double h = 0.5 * dt * kC_R;
    k.hx = (float)(h * C.idx); k.hy = (float)(h * C.idy); k.hz = (float)(h * C.idz);
    k.hk = (float)h;
    k.c6 = dt / 6.0 * kC_R;
    k.c6r = (float)k.c6;
    k.ix = (float)C.idx; k.iy = (float)C.idy; k.iz = (float)C.idz;
    k.perturb = (float)perturb_ratio;
    return k;
}

// Largest |cell offset| any stage of any of the three rays of a step can have; the FP32 stepper
// needs it < 1 (host-side dispatch, rtgrff_api.cu).
inline double max_stage_offset_cells(double dt, double perturb_ratio, double idx, double idy, double idz)
{
    const double i = fmax(idx, fmax(idy, idz));
    return (1.0 + fabs(perturb_ratio)) * dt * kC_R * i;
}

struct RkSum {
    float vx, vy, vz, gx, gy, gz;
};

// Classic RK4 (build_rays.py:177-182) in cell-relative FP32: returns sum = k1 + 2 k2 + 2 k3 + k4.
template <bool EDGE>
__device__ __forceinline__ RkSum rk4_32(const RayCube &C, const StepConst &K, int bi, int bj, int bk, int base_off,
                                        Cell &cache, float px, float py, float pz, float kx, float ky, float kz)
{
    const Deriv32 k1 = rhs32<EDGE>(C, bi, bj, bk, base_off, cache, px, py, pz, kx, ky, kz);
    const Deriv32 k2 = rhs32<EDGE>(C, bi, bj, bk, base_off, cache, fmaf(K.hx, k1.vx, px), fmaf(K.hy, k1.vy, py),
                                   fmaf(K.hz, k1.vz, pz), fmaf(-K.hk, k1.gx, kx), fmaf(-K.hk, k1.gy, ky),
                                   fmaf(-K.hk, k1.gz, kz));
    const Deriv32 k3 = rhs32<EDGE>(C, bi, bj, bk, base_off, cache, fmaf(K.hx, k2.vx, px), fmaf(K.hy, k2.vy, py),
                                   fmaf(K.hz, k2.vz, pz), fmaf(-K.hk, k2.gx, kx), fmaf(-K.hk, k2.gy, ky),
                                   fmaf(-K.hk, k2.gz, kz));
    const float fx = 2.0f * K.hx, fy = 2.0f * K.hy, fz = 2.0f * K.hz, fk = 2.0f * K.hk;
    const Deriv32 k4 = rhs32<EDGE>(C, bi, bj, bk, base_off, cache, fmaf(fx, k3.vx, px), fmaf(fy, k3.vy, py),
                                   fmaf(fz, k3.vz, pz), fmaf(-fk, k3.gx, kx), fmaf(-fk, k3.gy, ky),
                                   fmaf(-fk, k3.gz, kz));
    RkSum s;
    s.vx = fmaf(2.0f, k2.vx, k1.vx) + fmaf(2.0f, k3.vx, k4.vx);
    s.vy = fmaf(2.0f, k2.vy, k1.vy) + fmaf(2.0f, k3.vy, k4.vy);
    s.vz = fmaf(2.0f, k2.vz, k1.vz) + fmaf(2.0f, k3.vz, k4.vz);
    s.gx = fmaf(2.0f, k2.gx, k1.gx) + fmaf(2.0f, k3.gx, k4.gx);
    s.gy = fmaf(2.0f, k2.gy, k1.gy) + fmaf(2.0f, k3.gy, k4.gy);
    s.gz = fmaf(2.0f, k2.gz, k1.gz) + fmaf(2.0f, k3.gz, k4.gz);
    return s;
}

template <bool CS, bool EDGE>
__device__ __forceinline__ void step32_body(const RayCube &C, const StepConst &K, Cell &cache, State &s, int bi, int bj,
                                            int bk, float px, float py, float pz, bool want_s, double &s_step)
{
    const int base_off = (bi * C.ny + bj) * C.nz + bk;
    const float kx = (float)s.kx, ky = (float)s.ky, kz = (float)s.kz;
    const RkSum c = rk4_32<EDGE>(C, K, bi, bj, bk, base_off, cache, px, py, pz, kx, ky, kz);
    s.rx = fma((double)c.vx, K.c6, s.rx); s.ry = fma((double)c.vy, K.c6, s.ry); s.rz = fma((double)c.vz, K.c6, s.rz);
    s.kx = fma((double)c.gx, -K.c6, s.kx); s.ky = fma((double)c.gy, -K.c6, s.ky); s.kz = fma((double)c.gz, -K.c6, s.kz);
    if (CS && want_s) {
        // build_rays.py:209-239 with d = r_pert' - r_central' = eps*e + c6*(sum_v_pert - sum_v_central)
        const float dx = K.c6r * c.vx, dy = K.c6r * c.vy, dz = K.c6r * c.vz;
        const float nrd = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        const float inv = 1.0f / (nrd + 1e-32f);
        const float tx = dx * inv, ty = dy * inv, tz = dz * inv;
        const bool use_z = fabsf(tz) < 0.9f;
        float e1x = use_z ? -ty : tz, e1y = use_z ? tx : 0.0f, e1z = use_z ? 0.0f : -tx;
        const float n1 = 1.0f / (sqrtf(fmaf(e1x, e1x, fmaf(e1y, e1y, e1z * e1z))) + 1e-30f);
        e1x *= n1; e1y *= n1; e1z *= n1;
        float e2x = ty * e1z - tz * e1y, e2y = tz * e1x - tx * e1z, e2z = tx * e1y - ty * e1x;
        const float n2 = 1.0f / (sqrtf(fmaf(e2x, e2x, fmaf(e2y, e2y, e2z * e2z))) + 1e-30f);
        e2x *= n2; e2y *= n2; e2z *= n2;
        const float eps = K.perturb * nrd;
        float d1x = 0.f, d1y = 0.f, d1z = 0.f, d2x = 0.f, d2y = 0.f, d2z = 0.f;
        // the two pencil rays share one code body (keeps the kernel inside the instruction cache)
#ifdef RT_PENCIL_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int q = 0; q < 2; ++q) {
            const float ex = eps * (q ? e2x : e1x), ey = eps * (q ? e2y : e1y), ez = eps * (q ? e2z : e1z);
            const RkSum a = rk4_32<EDGE>(C, K, bi, bj, bk, base_off, cache, fmaf(ex, K.ix, px), fmaf(ey, K.iy, py),
                                         fmaf(ez, K.iz, pz), kx, ky, kz);
            const float ddx = fmaf(K.c6r, a.vx - c.vx, ex), ddy = fmaf(K.c6r, a.vy - c.vy, ey),
                        ddz = fmaf(K.c6r, a.vz - c.vz, ez);
            if (q) { d2x = ddx; d2y = ddy; d2z = ddz; } else { d1x = ddx; d1y = ddy; d1z = ddz; }
        }
        const float cx = d1y * d2z - d1z * d2y, cy = d1z * d2x - d1x * d2z, cz = d1x * d2y - d1y * d2x;
        s_step = (double)(fabsf(fmaf(cx, tx, fmaf(cy, ty, cz * tz))) / (eps * eps));
    }
}

// One full step of the master state `s` (which must be inside the cube): central RK4 and, when
// `want_s` (warp-uniform), the cross-section ratio of this step.  Returns true if the state changed.
template <bool CS>
__device__ __forceinline__ bool step32(const RayCube &C, const StepConst &K, Cell &cache, State &s, bool want_s,
                                       double &s_step)
{
    // split the master position into base cell + fraction (FP64 -> FP32 once per step)
    const double fx = (s.rx - C.x0) * C.idx, fy = (s.ry - C.y0) * C.idy, fz = (s.rz - C.z0) * C.idz;
    const int bi = min((int)fx, C.nx - 2), bj = min((int)fy, C.ny - 2), bk = min((int)fz, C.nz - 2);
    const float px = (float)(fx - (double)bi), py = (float)(fy - (double)bj), pz = (float)(fz - (double)bk);
    const State s0 = s;
    // stages and pencil rays stay within one cell of the base cell: only a base cell next to a face
    // can produce an out-of-cube stage
    const bool edge = (bi < 1) | (bi > C.nx - 3) | (bj < 1) | (bj > C.ny - 3) | (bk < 1) | (bk > C.nz - 3);
    if (edge) step32_body<CS, true>(C, K, cache, s, bi, bj, bk, px, py, pz, want_s, s_step);
    else step32_body<CS, false>(C, K, cache, s, bi, bj, bk, px, py, pz, want_s, s_step);
    return state_differs(s, s0);
}

}  // namespace rtgrff
