"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU restatement in ``oracle/*.c`` plus numpy restatements of the
reference's host-side glue.  Nothing under ``raytracinggrff_b200/`` imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and only as the checker / CPU baseline.

Function -> reference location (all paths under /root/reference):
  ray_trace                    raytracingGRFF/build_rays.py:128-248
  sample_model_with_rays_cpu   raytracingGRFF/gpu_raytrace.py:632-651 (+ :473-535)
  check_uniform_grid           raytracingGRFF/gpu_raytrace.py:21-33
  get_mw / get_mw_slice        GRFF PyGET_MW / fastGRFF get_mw_slice call sites,
                               script/resample_with_ray_tracing.py:79-86, :404-446, :502-509
                               (PARITY UNPINNED: GRFF is an absent third-party binary)
  ray_launch_geometry          script/resample_with_ray_tracing.py:295-303
  emission_from_samples        script/resample_with_ray_tracing.py:354-365, :467-530
  emission_from_los            script/synthetic_FF_map_single_thread.py:149-224
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

C_R = 2.998e10 / 6.96e10  # build_rays.py:29-32

# workflow constants, script/resample_with_ray_tracing.py:68, :91-94
R_SUN_CM = 6.957e10
C_CGS = 2.998e10
KB_CGS = 1.38065e-16
SFU2CGS = 1e-19
AU_CM = 1.49599e13


def build(force: bool = False) -> Path:
    """Compile oracle/liboracle.so with gcc (Makefile in this directory)."""
    so = _HERE / "liboracle.so"
    srcs = sorted(_HERE.glob("oracle_*.c")) + [_HERE / "Makefile"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        dp = ctypes.POINTER(ctypes.c_double)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        _LIB.oracle_ray_trace.argtypes = [dp, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, dp, dp, dp, dp, ctypes.c_long,
                                          ctypes.c_double, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_int, dp, dp,
                                          ctypes.POINTER(ctypes.c_longlong)]
        _LIB.oracle_ray_trace.restype = ctypes.c_int
        _LIB.oracle_gradient.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_int, dp]
        _LIB.oracle_gradient.restype = None
        _LIB.oracle_sample_model.argtypes = [fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int] + \
            [ctypes.c_double] * 6 + [fp, fp, fp, ctypes.c_long, ctypes.c_long] + [ctypes.c_double] * 4 + \
            [fp, fp, fp, fp, ctypes.POINTER(ctypes.c_uint8)]
        _LIB.oracle_sample_model.restype = ctypes.c_int
        _LIB.oracle_get_mw.argtypes = [ip, dp, dp, dp, dp, dp, dp]
        _LIB.oracle_get_mw.restype = ctypes.c_int
        _LIB.oracle_get_mw_slice.argtypes = [ip, dp, dp, dp, dp, dp, dp, ip]
        _LIB.oracle_get_mw_slice.restype = ctypes.c_int
        _LIB.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _LIB.oracle_set_num_threads.restype = None
    return _LIB


def set_num_threads(n=None):
    """OpenMP threads of every oracle stage (default: all host cores, whatever OMP_NUM_THREADS says —
    torchrun sets it to 1 for its workers)."""
    n = int(n or os.cpu_count() or 1)
    _lib().oracle_set_num_threads(n)
    return n


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def check_uniform_grid(grid, name):
    """gpu_raytrace.py:21-33."""
    g = np.asarray(grid, dtype=np.float64)
    if g.ndim != 1 or g.size < 2:
        raise ValueError(f"{name} must be 1D with at least 2 points")
    d = np.diff(g)
    step = float(np.mean(d))
    if not np.isfinite(step) or step <= 0.0:
        raise ValueError(f"{name} has invalid spacing")
    max_dev = float(np.max(np.abs(d - step)))
    tol = max(1e-6 * abs(step), 1e-7 * max(abs(g[0]), abs(g[-1]), 1.0))
    if max_dev > tol:
        raise ValueError(f"{name} must be uniformly spaced")
    return float(g[0]), step


def gradient(f, h, axis):
    f = _f64(f)
    out = np.empty_like(f)
    _lib().oracle_gradient(_p(f, ctypes.c_double), *f.shape, float(h), int(axis), _p(out, ctypes.c_double))
    return out


def ray_trace(omega_pe_3d, x_grid, y_grid, z_grid, freq_hz, x_start, y_start, z_start,
              kvec_in_norm, dt, n_steps, record_stride=10, trace_crosssections=False,
              cross_section_stride=1, perturb_ratio=2, n_threads=0, return_active=False, cache_gradient=False):
    """Same signature and return convention as build_rays.ray_trace (build_rays.py:128-130, :248):
    (r_record float64 (n_rec,n_rays,3), list of n_rec float64 (n_rays,) arrays or []).
    cache_gradient: keep np.gradient(omega_pe) between calls on the same cube array (the benchmark's
    sub-sampled CPU legs: over a full map that preparation is amortised over millions of rays; the caller
    must not modify the array in between)."""
    w = _f64(omega_pe_3d)
    xg, yg, zg = _f64(x_grid), _f64(y_grid), _f64(z_grid)
    xs, ys, zs = _f64(x_start), _f64(y_start), _f64(z_start)
    kv = _f64(kvec_in_norm)
    n_rays = xs.shape[0]
    n_steps, stride = int(n_steps), int(record_stride)
    n_rec = (n_steps + stride - 1) // stride if n_steps > 0 else 0
    r_record = np.empty((n_rec, n_rays, 3), dtype=np.float64)
    s_record = np.empty((n_rec, n_rays), dtype=np.float64)
    active = ctypes.c_longlong(0)
    d = ctypes.c_double
    rc = _lib().oracle_ray_trace(_p(w, d), _p(xg, d), _p(yg, d), _p(zg, d), *w.shape, float(freq_hz),
                                 _p(xs, d), _p(ys, d), _p(zs, d), _p(kv, d), n_rays, float(dt), n_steps,
                                 stride, int(bool(trace_crosssections)), float(perturb_ratio),
                                 int(n_threads) if not cache_gradient else -1 - int(n_threads),
                                 _p(r_record, d), _p(s_record, d), ctypes.byref(active))
    if rc != 0:
        raise MemoryError("oracle_ray_trace allocation failed")
    cs = [s_record[i].copy() for i in range(n_rec)] if trace_crosssections else []
    if return_active:
        return r_record, cs, int(active.value)
    return r_record, cs


def sample_model_with_rays_cpu(x_grid, y_grid, z_grid, ne_xyz, te_xyz, b_xyz, r_record, s_arr,
                               ray_start, r_sun_cm, fill_ne=0.0, fill_te=1e4, fill_b=0.0):
    """gpu_raytrace.py:632-651; returns the same dict."""
    x0, dx = check_uniform_grid(np.asarray(x_grid), "x_grid")
    y0, dy = check_uniform_grid(np.asarray(y_grid), "y_grid")
    z0, dz = check_uniform_grid(np.asarray(z_grid), "z_grid")
    pos = _f32(np.asarray(r_record))
    s = _f32(np.asarray(s_arr))
    rs = _f32(ray_start)
    ne3, te3, b3 = _f32(ne_xyz), _f32(te_xyz), _f32(b_xyz)
    n_rec, n_rays, _ = pos.shape
    out = {k: np.empty((n_rec, n_rays), dtype=np.float32) for k in ("ne", "te", "b", "ds")}
    valid = np.empty((n_rec, n_rays), dtype=np.uint8)
    f = ctypes.c_float
    _lib().oracle_sample_model(_p(ne3, f), _p(te3, f), _p(b3, f), *ne3.shape, x0, dx, y0, dy, z0, dz,
                               _p(pos, f), _p(s, f), _p(rs, f), n_rec, n_rays, float(r_sun_cm),
                               float(fill_ne), float(fill_te), float(fill_b),
                               _p(out["ne"], f), _p(out["te"], f), _p(out["b"], f), _p(out["ds"], f),
                               _p(valid, ctypes.c_uint8))
    out["valid_mask"] = valid.astype(bool)
    out["s"] = s
    return out


def get_mw(Lparms, Rparms, Parms, T_arr, DEM_arr, DDM_arr, RL):
    """PyGET_MW contract (script/resample_with_ray_tracing.py:79-86); writes RL in place."""
    L = np.asfortranarray(Lparms, dtype=np.int32)
    R = np.asfortranarray(Rparms, dtype=np.float64)
    P = np.asfortranarray(Parms, dtype=np.float64)
    assert RL.flags.f_contiguous and RL.dtype == np.float64
    d = ctypes.c_double
    dummy = np.zeros(1)
    return _lib().oracle_get_mw(_p(L, ctypes.c_int32), _p(R, d), _p(P, d), _p(dummy, d), _p(dummy, d),
                                _p(dummy, d), _p(RL, d))


def get_mw_slice(Lparms_M, Rparms_M, Parms_M, T_arr, DEM_arr, DDM_arr, RL_M):
    """fastGRFF get_mw_slice contract (script/resample_with_ray_tracing.py:428-446); returns status."""
    L = np.asfortranarray(Lparms_M, dtype=np.int32)
    R = np.asfortranarray(Rparms_M, dtype=np.float64)
    P = np.asfortranarray(Parms_M, dtype=np.float64)
    assert RL_M.flags.f_contiguous and RL_M.dtype == np.float64
    status = np.zeros(int(L[0]), dtype=np.int32)
    d = ctypes.c_double
    dummy = np.zeros(1)
    rc = _lib().oracle_get_mw_slice(_p(L, ctypes.c_int32), _p(R, d), _p(P, d), _p(dummy, d), _p(dummy, d),
                                    _p(dummy, d), _p(RL_M, d), _p(status, ctypes.c_int32))
    if rc != 0:
        raise ValueError("oracle_get_mw_slice: bad sizes")
    return status


def ray_launch_geometry(N_pix, X_fov, z_observer):
    """script/resample_with_ray_tracing.py:295-303."""
    x_coords = np.linspace(-X_fov, X_fov, N_pix)
    y_coords = np.linspace(-X_fov, X_fov, N_pix)
    X_img, Y_img = np.meshgrid(x_coords, y_coords)
    x_flat = X_img.ravel()
    y_flat = Y_img.ravel()
    z_start = np.sqrt(np.abs((z_observer * 2.0) ** 2 - x_flat ** 2 - y_flat ** 2)) / 2.0
    kvec = np.tile([[0, 0, -1]], (len(x_flat), 1))
    return x_flat, y_flat, z_start, kvec


def pack_parms_batch(sampled, pixel_area_cm2, s_input_on=False):
    """script/resample_with_ray_tracing.py:404-426: Parms_M (15, n_rec, n_rays) Fortran order,
    valid samples compacted to the front in record order, zero padding after."""
    ne_all, te_all, b_all = sampled["ne"], sampled["te"], sampled["b"]
    ds_all, valid_all, s_all = sampled["ds"], sampled["valid_mask"], sampled["s"]
    n_rec, n_rays = ne_all.shape
    Parms_M = np.zeros((15, n_rec, n_rays), dtype=np.float64, order="F")
    Parms_M[4, :, :] = 90.0
    Parms_M[6, :, :] = 1 + 4
    Parms_M[7, :, :] = 30
    for p in range(n_rays):
        valid = valid_all[:, p] & np.isfinite(ne_all[:, p]) & np.isfinite(te_all[:, p]) & np.isfinite(b_all[:, p])
        if not np.any(valid):
            continue
        cnt = int(np.count_nonzero(valid))
        Parms_M[0, :cnt, p] = ds_all[:, p][valid]
        Parms_M[1, :cnt, p] = te_all[:, p][valid]
        Parms_M[2, :cnt, p] = ne_all[:, p][valid]
        Parms_M[3, :cnt, p] = b_all[:, p][valid]
        Parms_M[14, :cnt, p] = s_all[:, p][valid] * pixel_area_cm2 if s_input_on else 0.0
    return Parms_M


def emission_from_samples(sampled, N_pix, X_fov, freq0, Nfreq=1, freq_log_step=0.0, s_input_on=False):
    """script/resample_with_ray_tracing.py:354-365 (Lparms/Rparms), :467-524 (per-pixel GET_MW loop),
    :530 (scrub).  Returns (emission_cube, emission_polVI_cube, frequencies_Hz)."""
    Nf = int(Nfreq)
    frequencies_Hz = freq0 * (10.0 ** (freq_log_step * np.arange(Nf)))
    Lparms = np.zeros(5, dtype="int32")
    Lparms[1] = Nf
    Rparms = np.zeros(3, dtype="double")
    pixel_size_cm = (2 * X_fov) / N_pix * R_SUN_CM
    pixel_area_cm2 = pixel_size_cm * pixel_size_cm
    Rparms[0] = pixel_area_cm2
    Rparms[1] = freq0
    Rparms[2] = freq_log_step
    emission_cube = np.zeros((N_pix, N_pix, Nf), dtype="double")
    emission_polVI_cube = np.zeros((N_pix, N_pix, Nf), dtype="double")
    ne_all, te_all, b_all = sampled["ne"], sampled["te"], sampled["b"]
    ds_all, valid_all, s_all = sampled["ds"], sampled["valid_mask"], sampled["s"]
    n_rays = ne_all.shape[1]
    for p in range(n_rays):
        i, j = p // N_pix, p % N_pix
        valid = valid_all[:, p] & np.isfinite(ne_all[:, p]) & np.isfinite(te_all[:, p]) & np.isfinite(b_all[:, p])
        if not np.any(valid):
            continue
        n_valid = int(np.count_nonzero(valid))
        Parms = np.zeros((15, n_valid), dtype="double", order="F")
        Parms[0] = ds_all[:, p][valid]
        Parms[1] = te_all[:, p][valid]
        Parms[2] = ne_all[:, p][valid]
        Parms[3] = b_all[:, p][valid]
        Parms[4] = 90.0
        Parms[6] = 1 + 4
        Parms[7] = 30
        Parms[14] = s_all[:, p][valid] * pixel_area_cm2 if s_input_on else 0.0
        L = Lparms.copy()
        L[0] = n_valid
        RL = np.zeros((7, Nf), dtype="double", order="F")
        if get_mw(L, Rparms, Parms, None, None, None, RL) != 0:
            continue
        for f in range(Nf):
            inten = RL[5, f] + RL[6, f]
            pol = (RL[5, f] - RL[6, f]) / (RL[5, f] + RL[6, f] + 1e-30)
            nu = frequencies_Hz[f] if RL[0, f] <= 0 else RL[0, f] * 1e9
            conv = (SFU2CGS * C_CGS * C_CGS / (2.0 * KB_CGS * nu * nu) / Rparms[0]) * (AU_CM * AU_CM)
            emission_cube[i, j, f] = inten * conv
            emission_polVI_cube[i, j, f] = pol
    emission_cube = np.nan_to_num(emission_cube, nan=0.0, posinf=0.0, neginf=0.0)
    return emission_cube, emission_polVI_cube, frequencies_Hz


def chain_bvec(cube, freq_hz, dt, n_steps, record_stride, xs, ys, zs, area, em_flag=4, s_max=30, perturb_ratio=2,
               n_threads=0, ray_chunk=128, return_paths=False, cache_gradient=False):
    """The reference chain for a GR+FF map with the angle to B taken along the ray — the physics the fused
    kernel is benchmarked on (BASELINE configs 4 and 5), restated with the reference's own stages:
      ray_trace (build_rays.py:128-248, cross-sections on)
      -> sample_model_with_rays twice, (ne, te, b) and (bx, by, bz)     (gpu_raytrace.py:632-651)
      -> Parms per pixel as script/resample_with_ray_tracing.py:472-501, except Parms[3] = |B vector| and
         Parms[4] = theta = acos(-B.d / |B||d|), d the step from the previous valid float32 sample (the ray
         start for the first) to this one; Parms[6] = em_flag, Parms[7] = s_max
      -> GET_MW (batched over pixels) -> T_b, V/I (:513-520).
    The reference itself fixes theta = 90 deg (:495); theta from B is this repo's extension.
    cube: dict with x_grid, y_grid, z_grid, omega_pe, ne, te, b, bx, by, bz.  Returns (tb, vi) each (n_rays,)
    [and r_record, s_record when return_paths]."""
    kv = np.tile([[0.0, 0.0, -1.0]], (len(xs), 1))
    ray_start = np.column_stack([xs, ys, zs])
    g3 = (cube["x_grid"], cube["y_grid"], cube["z_grid"])
    r, cs = ray_trace(cube["omega_pe"], *g3, freq_hz, xs, ys, zs, kv, dt, n_steps, record_stride, True,
                      perturb_ratio=perturb_ratio, n_threads=n_threads, cache_gradient=cache_gradient)
    s_rec = np.array(cs)
    smp = sample_model_with_rays_cpu(*g3, cube["ne"], cube["te"], cube["b"], r, s_rec, ray_start, R_SUN_CM)
    bv = sample_model_with_rays_cpu(*g3, cube["bx"], cube["by"], cube["bz"], r, s_rec, ray_start, R_SUN_CM,
                                    fill_ne=0.0, fill_te=0.0, fill_b=0.0)
    tb, vi = emission_bvec_from_samples(smp, bv, r, ray_start, area, freq_hz, em_flag, s_max, ray_chunk)
    if return_paths:
        return tb, vi, r, s_rec
    return tb, vi


def emission_bvec_from_samples(smp, bv, r_record, ray_start, area, freq_hz, em_flag=4, s_max=30, ray_chunk=128):
    """Parms packing with theta from the sampled B vector + batched GET_MW + T_b conversion (see chain_bvec)."""
    n_rec, n_rays = smp["ne"].shape
    tb = np.zeros(n_rays)
    vi = np.zeros(n_rays)
    conv = (SFU2CGS * C_CGS * C_CGS / (2.0 * KB_CGS * freq_hz * freq_hz) / area) * (AU_CM * AU_CM)
    for c0 in range(0, n_rays, ray_chunk):
        sl = slice(c0, min(n_rays, c0 + ray_chunk))
        nr = sl.stop - sl.start
        valid = smp["valid_mask"][:, sl]
        pos = r_record[:, sl].astype(np.float32).astype(np.float64)
        start = np.asarray(ray_start)[sl].astype(np.float32).astype(np.float64)
        # previous valid sample of every record (the ray start before the first one)
        idx = np.where(valid, np.arange(n_rec)[:, None], -1)
        last = np.maximum.accumulate(idx, axis=0)
        prev = np.vstack([np.full((1, nr), -1, dtype=last.dtype), last[:-1]])
        prev_pos = np.take_along_axis(pos, np.broadcast_to(np.maximum(prev, 0)[:, :, None], pos.shape), axis=0)
        prev_pos = np.where(prev[:, :, None] >= 0, prev_pos, start[None])
        d = pos - prev_pos
        B = np.stack([bv["ne"][:, sl], bv["te"][:, sl], bv["b"][:, sl]], axis=2).astype(np.float64)
        b2 = (B * B).sum(2)
        dn2 = (d * d).sum(2)
        with np.errstate(divide="ignore", invalid="ignore"):
            cth = np.clip(-(B * d).sum(2) / np.sqrt(b2 * dn2), -1.0, 1.0)
        cth = np.where((b2 > 0) & (dn2 > 0), cth, 6.123233995736766e-17)
        theta = np.degrees(np.arccos(cth))
        ne, te, ds = smp["ne"][:, sl], smp["te"][:, sl], smp["ds"][:, sl]
        bmag = np.sqrt(b2)
        keep_src = valid & np.isfinite(ne) & np.isfinite(te) & np.isfinite(bmag)
        order = np.argsort(~keep_src, axis=0, kind="stable")
        cnt = keep_src.sum(axis=0)
        nz = int(cnt.max()) if nr else 0
        if nz == 0:
            continue
        keep = np.arange(nz)[:, None] < cnt[None, :]
        P = np.zeros((15, nz, nr), dtype=np.float64, order="F")
        for m, a in ((0, ds), (1, te), (2, ne), (3, bmag), (4, theta)):
            P[m] = np.where(keep, np.take_along_axis(a.astype(np.float64), order, axis=0)[:nz], 0.0)
        P[6] = em_flag
        P[7] = s_max
        L = np.array([nr, nz, 1, 1, 0, 0], dtype=np.int32)
        R = np.zeros((3, nr), order="F")
        R[0], R[1], R[2] = area, freq_hz, 0.0
        RL = np.zeros((7, 1, nr), order="F")
        status = get_mw_slice(L, R, P, None, None, None, RL)
        ok = (status == 0) & (cnt > 0)
        inten = RL[5, 0] + RL[6, 0]
        tb[sl] = np.where(ok, inten * conv, 0.0)
        vi[sl] = np.where(ok, (RL[5, 0] - RL[6, 0]) / (inten + 1e-30), 0.0)
    tb = np.nan_to_num(tb, nan=0.0, posinf=0.0, neginf=0.0)
    return tb, vi


def emission_from_los(Ne_LOS, Te_LOS, B_LOS, ds_LOS, pixel_area_cm2, freq0, Nfreq, freq_log_step):
    """script/synthetic_FF_map_single_thread.py:149-224 (straight-LOS GRFF map)."""
    N_pix = Ne_LOS.shape[0]
    Nf = int(Nfreq)
    frequencies_Hz = freq0 * (10.0 ** (freq_log_step * np.arange(Nf)))
    Rparms = np.array([pixel_area_cm2, freq0, freq_log_step], dtype="double")
    emission_cube = np.zeros((N_pix, Ne_LOS.shape[1], Nf), dtype="double")
    emission_polVI_cube = np.zeros_like(emission_cube)
    for i in range(N_pix):
        for j in range(Ne_LOS.shape[1]):
            ne, te, b, ds = Ne_LOS[i, j], Te_LOS[i, j], B_LOS[i, j], ds_LOS[i, j]
            vm = ~(np.isnan(ne) | np.isnan(te) | np.isnan(b))
            n_valid = int(np.count_nonzero(vm))
            if n_valid == 0:
                continue
            Parms = np.zeros((15, n_valid), dtype="double", order="F")
            Parms[0], Parms[1], Parms[2], Parms[3] = ds[vm], te[vm], ne[vm], b[vm]
            Parms[4] = 90.0
            Parms[6] = 1 + 4
            Parms[7] = 30
            L = np.array([n_valid, Nf, 0, 0, 0], dtype="int32")
            RL = np.zeros((7, Nf), dtype="double", order="F")
            if get_mw(L, Rparms, Parms, None, None, None, RL) != 0:
                continue
            for f in range(Nf):
                inten = RL[5, f] + RL[6, f]
                with np.errstate(invalid="ignore", divide="ignore"):
                    pol = (RL[5, f] - RL[6, f]) / (RL[5, f] + RL[6, f])
                nu = frequencies_Hz[f] if RL[0, f] <= 0 else RL[0, f] * 1e9
                conv = (SFU2CGS * C_CGS * C_CGS / (2.0 * KB_CGS * nu * nu) / Rparms[0]) * (1.49599e13 ** 2)
                emission_cube[i, j, f] = inten * conv
                emission_polVI_cube[i, j, f] = pol
    return emission_cube, emission_polVI_cube, frequencies_Hz


# ---------------------------------------------------------------------------------------------
# Image-plane post-processing
# ---------------------------------------------------------------------------------------------
def gaussian_filter(a, sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter(a, sigma) for a 2-D float64 map, restated in numpy: the call the
    workflow makes at script/resample_with_ray_tracing.py:618-624 and
    script/pub/compare_on_off_scaling_factor.py:51-69.  Separable; per axis (0 first) a symmetric
    correlation with weights exp(-x^2 / (2 sigma^2)) / sum over x = -r..r, r = int(truncate*sigma+0.5),
    'reflect' boundary (d c b a | a b c d | d c b a).  Summation order as scipy's symmetric
    correlate1d: centre tap, then the tap pairs from the farthest inwards."""
    out = np.array(a, dtype=np.float64, copy=True)
    r = int(truncate * float(sigma) + 0.5)
    if sigma > 0:
        x = np.arange(-r, r + 1)
        w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    else:
        w = np.ones(1)
    w = w / w.sum()
    for axis in (0, 1):
        n = out.shape[axis]
        src = np.moveaxis(out, axis, 0)
        idx = np.arange(-r, n + r)
        period = 2 * n
        idx = np.mod(idx, period)
        idx = np.where(idx < n, idx, period - 1 - idx)
        ext = src[idx]                                   # reflected line of length n + 2 r
        acc = ext[r:r + n] * w[r]
        for d in range(r, 0, -1):
            acc = acc + (ext[r - d:r - d + n] + ext[r + d:r + d + n]) * w[r - d]
        out = np.moveaxis(acc, 0, axis)
    return np.ascontiguousarray(out)


def _patch_plane(a, max_passes):
    ny, nx = a.shape
    for _ in range(max_passes):
        bad = ~np.isfinite(a)                            # state at the start of the sweep (util.py:48)
        if not bad.any():
            return
        fixed = 0
        for i, j in zip(*np.nonzero(bad)):               # row-major order, patched in place (util.py:52-74)
            vals = []
            for line, pos, step in ((a[i, :], j, -1), (a[i, :], j, 1), (a[:, j], i, -1), (a[:, j], i, 1)):
                q = pos + step                           # left, right, down, up: nearest finite pixel now
                while 0 <= q < line.size:
                    if np.isfinite(line[q]):
                        vals.append(line[q])
                        break
                    q += step
            if vals:
                a[i, j] = np.mean(vals)
                fixed += 1
        if fixed == 0:
            return


def patch_nan_emission_map(emission, max_passes=10):
    """raytracingGRFF/util.py:6-77 restated: non-finite pixels of a (ny,nx) map or of each [:, :, k]
    slice of a (ny,nx,nf) cube become the mean of the nearest finite pixels left/right/below/above."""
    out = np.array(emission, dtype=np.float64, copy=True)
    if out.ndim == 2:
        _patch_plane(out, max_passes)
    elif out.ndim == 3:
        for k in range(out.shape[2]):
            _patch_plane(out[:, :, k], max_passes)
    else:
        raise ValueError("emission must be 2D or 3D")
    return out
