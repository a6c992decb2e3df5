"""Image-plane post-processing of the emission maps on the GPU.

Drop-ins for ``raytracingGRFF/util.py:6-77`` (``patch_nan_emission_map``, exported by the
reference package, ``__init__.py:3-15``) and for the Gaussian beam the workflow convolves its maps
with (``scipy.ndimage.gaussian_filter`` at ``script/resample_with_ray_tracing.py:618-624`` and
``script/pub/compare_on_off_scaling_factor.py:51-69``).  Both run in librtgrff_b200.so; there is
no CPU path here.
"""
from __future__ import annotations

from ctypes import c_double, c_int64, byref

import numpy as np

from . import _lib

# the workflow's constants (script/pub/compare_on_off_scaling_factor.py:32-34)
R_SUN_M = 6.957e8
AU_M = 1.495978707e11
C_M_S = 2.99792458e8


def _planes(a):
    """(ny, nx) or (ny, nx, nf) -> a fresh contiguous float64 (nf, ny, nx) copy."""
    if a.ndim == 2:
        return np.array(a[None], dtype=np.float64, order="C", copy=True)
    if a.ndim == 3:
        return np.array(np.moveaxis(a, 2, 0), dtype=np.float64, order="C", copy=True)
    raise ValueError("emission must be 2D or 3D")


def patch_nan_emission_map(emission, inplace=False, max_passes=10, device=0):
    """Fill non-finite pixels with the mean of the nearest finite pixels to the left, right, below
    and above (``raytracingGRFF/util.py:6-41``; 3-D input is patched per ``[:, :, k]`` slice).
    Same signature and result as the reference; ``max_passes`` is its ``_patch_nan_2d`` default."""
    src = np.asarray(emission)
    if src.ndim not in (2, 3):
        raise ValueError("emission must be 2D or 3D")
    planes = _planes(src)
    nf, ny, nx = planes.shape
    n = c_int64(0)
    if planes.size:
        ctx = _lib.default_context(device)
        _lib.check(ctx._lib.rtgrff_patch_nan(ctx.handle, _lib.ptr(planes, c_double), ny, nx, nf, int(max_passes), byref(n)))
    res = planes[0] if src.ndim == 2 else np.moveaxis(planes, 0, 2)
    if inplace:
        emission[...] = res
        return emission
    return np.array(res, dtype=np.float64, copy=True)


def gaussian_beam(emission_map, sigma, truncate=4.0, device=0):
    """``scipy.ndimage.gaussian_filter(emission_map, sigma=sigma)`` on the GPU for a 2-D map or, per
    frequency slice, a (ny, nx, nf) cube: float64, 'reflect' boundary, radius int(truncate*sigma+0.5)."""
    src = np.asarray(emission_map)
    planes = _planes(src)
    nf, ny, nx = planes.shape
    out = np.empty_like(planes)
    if planes.size:
        ctx = _lib.default_context(device)
        _lib.check(ctx._lib.rtgrff_gaussian_beam(ctx.handle, _lib.ptr(planes, c_double), ny, nx, nf, float(sigma),
                                                 float(truncate), _lib.ptr(out, c_double)))
    return out[0] if src.ndim == 2 else np.moveaxis(out, 0, 2)


def convolve_beam(emission_map, beam_fwhm, x_range, N_pix, device=0):
    """The ``--consider-beam`` step of the workflow (``script/resample_with_ray_tracing.py:618-624``):
    sigma [pixels] = beam_fwhm / (x_range[1] - x_range[0]) * N_pix — the reference passes this
    FWHM-derived number to ``gaussian_filter`` as sigma unchanged, and so does this function."""
    beam_radius_pix = beam_fwhm / (x_range[-1] - x_range[0]) * N_pix
    return gaussian_beam(emission_map, beam_radius_pix, device=device)


def apply_baseline_beam(tb_map, x_coords_m, y_coords_m, freq_hz, baseline_km, device=0):
    """``_apply_baseline_beam`` (``script/pub/compare_on_off_scaling_factor.py:51-69``): diffraction
    beam lambda / baseline at 1 au, FWHM -> sigma = FWHM / 2.355 pixels."""
    out = np.array(tb_map, dtype=float, copy=True)
    if baseline_km <= 0 or len(x_coords_m) < 2 or len(y_coords_m) < 2:
        return out
    pix_rsun = 0.5 * (abs((x_coords_m[1] - x_coords_m[0]) / R_SUN_M) + abs((y_coords_m[1] - y_coords_m[0]) / R_SUN_M))
    if pix_rsun <= 0:
        return out
    beam_fwhm_rsun = (C_M_S / freq_hz) / (baseline_km * 1e3) * AU_M / R_SUN_M
    sigma_pix = beam_fwhm_rsun / pix_rsun / 2.355
    if sigma_pix <= 0:
        return out
    return gaussian_beam(out, sigma_pix, device=device)


def dlogS_ds(r_record, s_record, s_floor=0.01, distance_ray=None):
    """Differential magnification d(log S)/ds [1/R_sun] along the rays, the diagnostic of
    ``script/pub/cross_section_plots.ipynb`` cell 12: S below ``s_floor`` -> NaN, first differences
    of log S over the path length between records, NaN/inf -> 0.  ``distance_ray`` = i reproduces the
    notebook, which divides every ray by the path length of ray i (it uses ray 0); the default
    (None) uses each ray's own path length.  Host-side numpy on traced records (r_record (R,N,3),
    s_record (R,N) as ``trace_ray`` returns them) — diagnostics, not part of the per-ray hot path."""
    r = np.asarray(r_record, dtype=np.float64)
    S = np.array(s_record, dtype=np.float64, copy=True)
    S[S < s_floor] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        dlog = np.diff(np.log(S), axis=0)
        seg = np.sqrt(np.sum(np.diff(r, axis=0) ** 2, axis=2))          # (R-1, N)
        if distance_ray is not None:
            seg = np.repeat(seg[:, [distance_ray]], S.shape[1], axis=1)
        out = dlog / seg
    out[~np.isfinite(out)] = 0.0
    return out
