"""ORACLE — TEST INFRASTRUCTURE ONLY.

Parity of a GPU-rendered GR+FF map against the CPU oracle chain on a pixel sub-sample: the numbers
BASELINE.md §4.4 asks every benchmark report to carry (max |dr| in R_sun, max relative dT_b, max |d(V/I)|)
and the parity tests on the benchmarked configurations assert on.  Used by ``tests/`` and by ``bench.py``
(outside its timed regions); nothing under ``raytracinggrff_b200/`` imports it.

Tolerances (BASELINE.json north_star): paths <= 1e-5 R_sun, T_b and V/I <= 1e-4.

Rays that cross the r = 1 density discontinuity of the model (n_e = 0 inside, script/
resample_with_ray_tracing.py:269-279) at grazing incidence are chaotic: they amplify the 1e-9 relative
difference between the float64 cube of the oracle and its float32 storage on the device to O(1) R_sun
(tests/test_gpu_parity.py::test_rays_through_the_density_discontinuity).  Those pixels ("diving": the
oracle's ray comes within one cell of r = 1 while inside the cube) are counted and reported separately;
every other pixel is held to the tolerances.
"""
from __future__ import annotations

import numpy as np

from . import oracle

POS_TOL = 1e-5
TB_RTOL = 1e-4
VI_ATOL = 1e-4


def subsample(n_x, n_y, stride, offset=None):
    """Flat indices p = i*n_x + j of every `stride`-th pixel in x and y."""
    o = stride // 2 if offset is None else offset
    return (np.arange(o, n_y, stride)[:, None] * n_x + np.arange(o, n_x, stride)[None, :]).ravel()


def map_parity(cube, freq_params, xs, ys, zs, area, tb_gpu, vi_gpu, session=None, em_flag=4, s_max=30,
               n_threads=None, cell=None):
    """cube: host dict (x_grid, y_grid, z_grid, omega_pe, ne, te, b, bx, by, bz) — the very arrays the
    device cubes were built from.  xs, ys, zs: the sub-sampled ray starts; tb_gpu, vi_gpu: (n_freq, n_rays)
    from the GPU map at those pixels.  session: a RaySession holding the same cubes; when given the rays
    are also traced on the GPU and the paths compared (max_dr_rsun).  Returns the parity dict."""
    oracle.set_num_threads(n_threads)
    lo = np.array([cube["x_grid"][0], cube["y_grid"][0], cube["z_grid"][0]])
    hi = np.array([cube["x_grid"][-1], cube["y_grid"][-1], cube["z_grid"][-1]])
    if cell is None:
        cell = float(cube["x_grid"][1] - cube["x_grid"][0])
    kv = np.tile([[0.0, 0.0, -1.0]], (len(xs), 1))
    out = dict(n_pixels=int(len(xs)), n_freq=len(freq_params), max_dr_rsun=0.0, max_rel_dTb=0.0, max_dVI=0.0,
               n_pixel_freqs_over_tol=0, n_diving_pixel_freqs=0, n_diving_over_tol=0, max_rel_dTb_diving=0.0,
               max_dr_rsun_diving=0.0, per_freq=[], oracle_seconds=0.0, oracle_nominal_ray_steps=0)
    import time
    for f, p in enumerate(freq_params):
        t0 = time.perf_counter()
        tb_ref, vi_ref, r_ref, _ = oracle.chain_bvec(cube, p["freq_hz"], p["dt"], p["n_steps"], p["record_stride"], xs,
                                                     ys, zs, area, em_flag, s_max, return_paths=True,
                                                     cache_gradient=True)
        out["oracle_seconds"] += time.perf_counter() - t0
        out["oracle_nominal_ray_steps"] += len(xs) * int(p["n_steps"])
        inside = np.all((r_ref >= lo) & (r_ref <= hi), axis=2)
        rad = np.where(inside, np.linalg.norm(r_ref, axis=2), np.inf)
        diving = rad.min(axis=0) < 1.0 + cell
        dr = np.zeros(len(xs))
        if session is not None:
            r_gpu, _, _ = session.trace(p["freq_hz"], xs, ys, zs, kv, p["dt"], p["n_steps"], p["record_stride"], True, 2.0)
            d = np.where(inside, np.abs(r_gpu - r_ref).max(axis=2), 0.0)
            dr = np.nan_to_num(d, nan=np.inf).max(axis=0)
            del r_gpu, d
        del r_ref
        tb, vi = np.asarray(tb_gpu[f], dtype=np.float64), np.asarray(vi_gpu[f], dtype=np.float64)
        zero_mismatch = (tb_ref == 0) != (tb == 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.where(tb_ref != 0, np.abs(tb - tb_ref) / np.abs(tb_ref), 0.0)
        rel = np.where(zero_mismatch, np.inf, rel)
        dvi = np.abs(vi - vi_ref)
        over = (rel > TB_RTOL) | (dvi > VI_ATOL) | (dr > POS_TOL)
        ok = ~diving
        row = dict(freq_hz=float(p["freq_hz"]), max_dr_rsun=float(dr[ok].max(initial=0.0)),
                   max_rel_dTb=float(rel[ok].max(initial=0.0)), max_dVI=float(dvi[ok].max(initial=0.0)),
                   n_over_tol=int((over & ok).sum()), n_diving=int(diving.sum()), n_diving_over_tol=int((over & diving).sum()),
                   median_tb=float(np.median(tb_ref[tb_ref > 0])) if (tb_ref > 0).any() else 0.0,
                   max_abs_vi=float(np.abs(vi_ref).max(initial=0.0)))
        out["per_freq"].append(row)
        out["max_dr_rsun"] = max(out["max_dr_rsun"], row["max_dr_rsun"])
        out["max_rel_dTb"] = max(out["max_rel_dTb"], row["max_rel_dTb"])
        out["max_dVI"] = max(out["max_dVI"], row["max_dVI"])
        out["n_pixel_freqs_over_tol"] += row["n_over_tol"]
        out["n_diving_pixel_freqs"] += row["n_diving"]
        out["n_diving_over_tol"] += row["n_diving_over_tol"]
        if diving.any():
            fin = np.isfinite(rel) & diving
            out["max_rel_dTb_diving"] = max(out["max_rel_dTb_diving"], float(rel[fin].max(initial=0.0)))
            out["max_dr_rsun_diving"] = max(out["max_dr_rsun_diving"],
                                            float(np.where(np.isfinite(dr), dr, 0.0)[diving].max(initial=0.0)))
    n = out["n_pixels"] * out["n_freq"]
    out["frac_pixels_over_tol"] = out["n_pixel_freqs_over_tol"] / max(n, 1)
    out["frac_diving"] = out["n_diving_pixel_freqs"] / max(n, 1)
    out["tolerances"] = dict(dr_rsun=POS_TOL, rel_dTb=TB_RTOL, dVI=VI_ATOL)
    out["reference"] = ("oracle chain: ray_trace -> sampler (ne,te,b | bx,by,bz) -> Parms with theta from B.d -> GET_MW; "
                        "pixels whose oracle ray comes within one cell of the r=1 density discontinuity are chaotic "
                        "and reported apart (diving)")
    return out
