// Fused map renderer for sm_100a: integrate the ray, sample n_e/T/|B| (and the B vector) at every
// recorded step and advance the polarised transfer equation — one thread per (pixel, frequency),
// nothing materialised: no r_record, no S array, no Parms.
//
// Equivalent to the reference chain
//   trace_ray / ray_trace                      (build_rays.py:128-248)
//   -> sample_model_with_rays                  (gpu_raytrace.py:632-651, float32 semantics)
//   -> Parms packing + GET_MW + T_b conversion (script/resample_with_ray_tracing.py:467-530)
// called once per frequency as the publication drivers do (script/pub/TbSpectra_gen.py:155-182).
// A 2048^2 x 500-record map would need 25 GB of paths per frequency if materialised
// (SURVEY.md §5); here a record lives in registers for the few hundred cycles it is needed.
#pragma once

#include "grff.cuh"
#include "los_sampler.cuh"
#include "trace_kernel.cuh"

#include <type_traits>

namespace rtgrff {

// Everything that depends on the frequency only, prepared on the host and passed BY VALUE inside
// the kernel parameters (constant bank, indexed by blockIdx.y): the ~20 registers these constants
// used to occupy per thread were spilled around the record-time code.
struct FreqDev {
    double nu, omega0, dt;
    int64_t n_steps, stride;
    int out_row;          // row of this frequency in the caller's tb / vi arrays
    StepConst K;
    FreqC fq;
};

// RT_FREQ_PER_LAUNCH = 1: one launch per frequency (on a few streams, so that the launches still overlap): the
// per-frequency constants then sit at fixed offsets of the constant bank and become instruction operands, where
// the multi-frequency launch indexes them with blockIdx.y (an LDC with a register index in the inner loops).
#ifndef RT_FREQ_PER_LAUNCH
#define RT_FREQ_PER_LAUNCH 1
#endif
constexpr int kMaxFreqPerLaunch = RT_FREQ_PER_LAUNCH;

#ifndef RT_PREFETCH
#define RT_PREFETCH 0     // 1: L1 prefetch of the sampler's lines one step ahead (A/B)
#endif

struct MapArgs {
    RayCube cube;
    const float4 *fcube, *bcube;
    GridGeomF fg;
    int64_t n_rays;
    const double *x_start, *y_start, *z_start, *kvec;
    const int *ray_order;                // thread t handles ray ray_order[t] (nullptr: t); see rtgrff.h
    int n_freq;                          // <= kMaxFreqPerLaunch
    FreqDev freqs[kMaxFreqPerLaunch];
    double perturb_ratio, area;
    float r_sun_cm, fill_ne, fill_te, fill_b;
    int em_flag, s_max, use_bvec, order, cs_every_step;
    int s_mode;                          // RTGRFF_S_PER_STEP / RTGRFF_S_CUMULATIVE: which S a record carries
    int s_input;                         // the record's S scales the voxel's source term (--s-input-on)
    int grff64;                          // 1: every voxel through the FP64 evaluation (RTGRFF_GRFF64=1, A/B and validation)
    double *tb, *vi;                     // [freq][ray]
    unsigned long long *active_steps;    // [0] active central steps, [1] steps with the pencil traced, [2] valid samples
};

// Transfer accumulated from the observer outwards (RTGRFF_ORDER_REVERSED): the records arrive
// nearest-first, the radiation travels farthest-first.  I_obs = acc + M I_far with M a 2x2 matrix
// in (L,R); folding one more (farther) operator I -> A I + b gives acc += M b, M = M A.
struct OutwardTransfer {
    double m00, m01, m10, m11, accL, accR;
    VoxLite prev;
    bool have_prev;
    __device__ __forceinline__ void init()
    {
        m00 = m11 = 1.0; m01 = m10 = 0.0; accL = accR = 0.0; have_prev = false;
    }
    __device__ __forceinline__ void fold(const DiagOp &d)
    {
        accL += m00 * d.bL + m01 * d.bR;
        accR += m10 * d.bL + m11 * d.bR;
        m00 *= d.aL; m10 *= d.aL; m01 *= d.aR; m11 *= d.aR;
    }
    __device__ __forceinline__ void fold_qt(double Q)
    {
        const double p = 1.0 - Q;
        const double a = m00 * Q + m01 * p, b = m00 * p + m01 * Q;
        const double c = m10 * Q + m11 * p, d = m10 * p + m11 * Q;
        m00 = a; m01 = b; m10 = c; m11 = d;
    }
};

// NEED_BETWEEN: gyroresonance on or theta from the B vector -> the previous voxel must be kept for
// the between-voxel events; with the reference's packing (theta = 90, GR off) it is dead weight.
// A voxel arrives as the float32 values the sampler produced (VoxLite + sin theta); its slab operator is
// evaluated in float32 where that is well conditioned (voxel_op_mixed), the intensities live in FP64.
template <bool NEED_BETWEEN>
struct RecordTransfer {
    PolState<1> st;
    VoxLite prev;
    bool have_prev;
    __device__ __forceinline__ void init() { st.clear(); have_prev = false; }
    __device__ __forceinline__ void push(const FreqC &f, const VoxLite &l, float sth, int flag, int smax, bool force64)
    {
        if (!voxel_nonempty_f(l.dz, l.T, l.ne, l.B, l.cth)) { have_prev = false; return; }
        if (NEED_BETWEEN) {
            if (have_prev && prev.B > 0.0f && l.B > 0.0f && between_needed_f(f, prev.cth, prev.B, l.cth, l.B, smax, !(flag & 1)))
            {
                Between b;
                between_voxels_cold(f.nu, f.sn, prev, l, flag, smax, &b);
                st.apply(b);
            }
            prev = l;
            have_prev = true;
        }
        st.apply(voxel_op_mixed(f, l.dz, l.T, l.ne, l.B, l.cth, sth, l.scale, flag, smax, force64));
    }
    __device__ __forceinline__ void result(double &L, double &R) const { L = st.L[0]; R = st.R[0]; }
};

template <bool NEED_BETWEEN>
struct OutwardTransferT : OutwardTransfer {
    __device__ __forceinline__ void push(const FreqC &f, const VoxLite &l, float sth, int flag, int smax, bool force64)
    {
        if (!voxel_nonempty_f(l.dz, l.T, l.ne, l.B, l.cth)) { have_prev = false; return; }
        if (NEED_BETWEEN) {
            if (have_prev && prev.B > 0.0f && l.B > 0.0f && between_needed_f(f, l.cth, l.B, prev.cth, prev.B, smax, !(flag & 1))) {
                // radiation crosses v -> prev: between_voxels(v, prev) lists the events in that order,
                // folding outwards meets them last-first
                Between b;
                between_voxels_cold(f.nu, f.sn, l, prev, flag, smax, &b);
                fold(b.after);
                if (b.qt) fold_qt(b.Q);
                fold(b.before);
            }
            prev = l;
            have_prev = true;
        }
        fold(voxel_op_mixed(f, l.dz, l.T, l.ne, l.B, l.cth, sth, l.scale, flag, smax, force64));
    }
    __device__ __forceinline__ void result(double &L, double &R) const { L = accL; R = accR; }
};

template <bool CS, int ORDER, bool BVEC, bool GR, int MODE>
__global__ void __launch_bounds__(RT_BLOCK, RT_MINB) render_map_kernel(const MapArgs a)
{
    constexpr bool NEED_BETWEEN = BVEC || GR;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int fi = (kMaxFreqPerLaunch == 1) ? 0 : (int)blockIdx.y;
    const bool has_ray = slot < a.n_rays;
    const int64_t ray = (has_ray && a.ray_order) ? (int64_t)a.ray_order[slot] : slot;
    const RayCube &C = a.cube;
    const FreqDev &fp = a.freqs[fi];
    const StepConst &K = fp.K;
    const FreqC &fq = fp.fq;
    Cell cache;
    cache.off = -1;

    State s;
    s.rx = s.ry = s.rz = s.kx = s.ky = s.kz = nan("");
    float px = 0.f, py = 0.f, pz = 0.f;   // previous VALID sample position, float32; starts at ray_start
    if (has_ray) {
        s.rx = a.x_start[ray]; s.ry = a.y_start[ray]; s.rz = a.z_start[ray];
        px = (float)s.rx; py = (float)s.ry; pz = (float)s.rz;     // _as_float32_c(ray_start), gpu_raytrace.py:650
        const double kc0 = start_kc(C, s.rx, s.ry, s.rz, fp.omega0);
        if (a.kvec) {
            s.kx = a.kvec[ray * 3 + 0] * kc0; s.ky = a.kvec[ray * 3 + 1] * kc0; s.kz = a.kvec[ray * 3 + 2] * kc0;
        } else {
            s.kx = 0.0 * kc0; s.ky = 0.0 * kc0; s.kz = -kc0;
        }
    }
    if (MODE == MODE_FAST32 && has_ray) init_cell(C, s, cache);
    bool alive = has_ray;
    float s_step = CS ? 0.0f : 1.0f;
    double s_cum = 1.0;
    const bool cumulative = CS && a.s_mode == RTGRFF_S_CUMULATIVE;
    unsigned int n_samples = 0;
    const int n_steps = (int)fp.n_steps, stride = (int)fp.stride;     // < 2^31, checked on the host
    int next_rec = 0;
    int death = -1;                       // step at which the ray froze (-1: still moving)

    typename std::conditional<ORDER == RTGRFF_ORDER_RECORD, RecordTransfer<NEED_BETWEEN>,
                              OutwardTransferT<NEED_BETWEEN>>::type tr;
    tr.init();
    bool first = true;
    bool tail_done = !has_ray;            // a frozen ray repeats the same record: ds = 0 -> empty voxel

    for (int i = 0; i < n_steps; ++i) {
#if RT_PREFETCH
        // the record taken after this step samples the field cubes in (or next to) the cell the ray is in now: ask L1
        // for those lines a whole step ahead (with a record at every step they are still there from the last one)
        if (stride > 1 && i == next_rec && alive) {
            int ci, cj, ck;
            float ftx, fty, ftz;
            if (cell_of(a.fg, (float)s.rx, (float)s.ry, (float)s.rz, ci, cj, ck, ftx, fty, ftz)) {
                const size_t sy = (size_t)a.fg.nz, sx = (size_t)a.fg.ny * a.fg.nz;
                const size_t off = (size_t)ci * sx + (size_t)cj * sy + (size_t)ck;
                const float4 *p0 = a.fcube + off;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p0));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p0 + sy));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p0 + sx));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p0 + sx + sy));
                if (BVEC) {
                    const float4 *p1 = a.bcube + off;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p1));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p1 + sy));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p1 + sx));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(p1 + sx + sy));
                }
            }
        }
#endif
        if (alive) {
            const bool want_s = CS && (i == next_rec || a.cs_every_step || cumulative);
            alive = advance_ray<CS, MODE>(C, K, cache, s, fp.dt, a.perturb_ratio, want_s, s_step);
            if (cumulative) s_cum *= s_step;     // gpu_raytrace.py:398-408
            if (!alive) death = i;
        }
        if (i == next_rec) {
            next_rec += stride;
            if (!tail_done) {
                // --- sampler (float32, gpu_raytrace.py:642-650) ---
                const float x = (float)s.rx, y = (float)s.ry, z = (float)s.rz, sv = cumulative ? (float)s_cum : s_step;
                if (sample_valid(x, y, z, sv)) {
                    ++n_samples;
                    float3 bv = make_float3(0.f, 0.f, 0.f);
                    FieldSample f = BVEC ? sample_fields_bvec_fast(a.fcube, a.bcube, a.fg, x, y, z, a.fill_ne, a.fill_te, bv)
                                         : sample_fields(a.fcube, a.fg, x, y, z, a.fill_ne, a.fill_te, a.fill_b);
                    const float dist = first ? dist_first_np(x, y, z, px, py, pz) : dist_fast(x, y, z, px, py, pz);
                    const float ds = __fmul_rn(dist, a.r_sun_cm);
                    float cthf = 6.123233995736766e-17f, sthf = 1.0f;                       // theta = 90 deg
                    if (BVEC) {
                        // theta between the B vector and the propagation direction (towards the observer:
                        // against the tracing direction), all from float32 inputs: FP32 throughout
                        const float dx = x - px, dy = y - py, dz = z - pz;
                        const float dn2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        const float b2 = fmaf(bv.x, bv.x, fmaf(bv.y, bv.y, bv.z * bv.z));
                        f.b = f.inb ? sqrt_approx(b2) : a.fill_b;
                        if (b2 > 0.0f && dn2 > 0.0f) {
                            const float inv = rsqrtf(b2 * dn2);
                            const float c = -fmaf(bv.x, dx, fmaf(bv.y, dy, bv.z * dz)) * inv;
                            cthf = fminf(1.0f, fmaxf(-1.0f, c));
                            // sin from the cross product: no cancellation near theta = 0
                            const float cx = bv.y * dz - bv.z * dy, cy = bv.z * dx - bv.x * dz, cz = bv.x * dy - bv.y * dx;
                            sthf = fminf(1.0f, sqrt_approx(fmaf(cx, cx, fmaf(cy, cy, cz * cz))) * inv);
                        }
                    }
                    // --- Parms packing rules (script/resample_with_ray_tracing.py:472-501) ---
                    if (isfinite(f.ne) && isfinite(f.te) && isfinite(f.b)) {
                        // Parms[14] = S * area (script/resample_with_ray_tracing.py:501): source factor S
                        const VoxLite lite = {ds, f.te, f.ne, f.b, cthf, a.s_input ? sv : 1.0f};
                        tr.push(fq, lite, sthf, a.em_flag, a.s_max, a.grff64 != 0);
                    }
                    px = x; py = y; pz = z;
                    first = false;
                }
                // once frozen, every later record repeats this one (ds = 0 or invalid): nothing to add
                if (!alive) tail_done = true;
            }
        }
        // a frozen ray still owes its first post-freeze record (with the cross-sections off its S is 1 and the
        // record counts, exactly as in trace -> sample -> emission): leave only when every lane has handed it over
        if (!__any_sync(0xffffffffu, alive || !tail_done)) break;
    }
    if (has_ray) {
        double tb, vi, IL, IR;
        tr.result(IL, IR);
        tb_vi(IL, IR, fp.nu, a.area, tb, vi);
        a.tb[(size_t)fp.out_row * a.n_rays + ray] = tb;
        a.vi[(size_t)fp.out_row * a.n_rays + ray] = vi;
    }
    if (a.active_steps) {
        // a ray moves on steps [0, death): the counters follow from where it froze
        unsigned long long moved_steps = has_ray ? (unsigned long long)(death >= 0 ? death : n_steps) : 0ull;
        const unsigned long long pencil_steps =
            !CS ? 0ull : ((a.cs_every_step || cumulative) ? moved_steps : (moved_steps + stride - 1) / stride);
        unsigned long long ps = pencil_steps, ns = n_samples;
        for (int off = 16; off > 0; off >>= 1) {
            moved_steps += __shfl_down_sync(0xffffffffu, moved_steps, off);
            ps += __shfl_down_sync(0xffffffffu, ps, off);
            ns += __shfl_down_sync(0xffffffffu, ns, off);
        }
        if ((threadIdx.x & 31) == 0) {
            if (moved_steps) atomicAdd(a.active_steps, moved_steps);
            if (ps) atomicAdd(a.active_steps + 1, ps);
            if (ns) atomicAdd(a.active_steps + 2, ns);
        }
    }
}

}  // namespace rtgrff
