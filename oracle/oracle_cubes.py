"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy/scipy restatement of the reference's cube preparation:
  cart_to_sph            raytracingGRFF/build_rays.py:35-45
  resample_to_xyz_cube   raytracingGRFF/build_rays.py:69-125 / script/resample_with_ray_tracing.py:110-151
  compose_cubes          script/resample_with_ray_tracing.py:263-293

PARITY UNPINNED for the sampler inside: the reference calls psipy's Variable.sample_at_coords,
and psipy is neither vendored nor installed here.  Its documented behaviour is restated — linear
interpolation (scipy interpn) on the (phi, latitude, r) mesh with the phi axis padded by one node
on each side to wrap — with one deliberate difference: a point outside the latitude / radius range
is NaN by itself (bounds_error=False) instead of psipy's exception voiding the whole x-slice
(build_rays.py:109-117).
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import RegularGridInterpolator


def cart_to_sph(x, y, z, phi0_offset=0.0):
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    colat = np.arccos(np.clip(z / r, -1.0, 1.0))
    lon = np.arctan2(y, x)
    lon = lon + phi0_offset * np.pi / 180.0
    lon = np.where(lon < 0, lon + 2 * np.pi, lon)
    return r, colat, lon


def sample_at_coords(var, lon, lat, r):
    """psipy Variable.sample_at_coords restated (phi padded to wrap, linear)."""
    phi = np.asarray(var.phi, dtype=np.float64)
    values = np.asarray(var.data)
    pc = np.concatenate([[phi[-1] - 2 * np.pi], phi, [phi[0] + 2 * np.pi]])
    vals = np.concatenate([values[-1:], values, values[:1]], axis=0)
    f = RegularGridInterpolator((pc, np.asarray(var.lat, dtype=np.float64), np.asarray(var.r, dtype=np.float64)),
                                vals, bounds_error=False, fill_value=np.nan)
    return f(np.column_stack([lon, lat, r])) * var.scale


def resample_to_xyz_cube(var, x_grid, y_grid, z_grid, phi0_offset=0.0, fill_nan=0.0, r_min=0.9999999):
    out = np.full((len(x_grid), len(y_grid), len(z_grid)), np.nan, dtype=float)
    y_mesh, z_mesh = np.meshgrid(y_grid, z_grid, indexing="ij")
    for ix, x_val in enumerate(x_grid):
        x_mesh = np.full_like(y_mesh, x_val)
        r, colat, lon = cart_to_sph(x_mesh, -z_mesh, y_mesh, phi0_offset=phi0_offset)
        lat = np.pi / 2 - colat
        r_mask = np.isfinite(r) & (r >= r_min)
        vals = np.full_like(r, np.nan, dtype=float)
        if np.any(r_mask):
            vals[r_mask] = sample_at_coords(var, lon[r_mask], lat[r_mask], r[r_mask])
        out[ix] = vals
    if fill_nan is not None:
        out = np.where(np.isfinite(out), out, fill_nan)
    return out


def compose_cubes(model, x_grid, y_grid, z_grid, phi0_offset=0.0, r_min=0.999999):
    temp = "te" if "te" in model else "t"
    rho = resample_to_xyz_cube(model["rho"], x_grid, y_grid, z_grid, phi0_offset, 0.0, r_min)
    omega_pe = 8.93e3 * np.sqrt(np.maximum(rho, 0.0)) * 2 * np.pi
    omega_pe = np.nan_to_num(omega_pe, nan=0.0, posinf=0.0, neginf=0.0)
    ne = np.maximum(rho, 0.0)
    te = resample_to_xyz_cube(model[temp], x_grid, y_grid, z_grid, phi0_offset, None, r_min)
    te = np.where(np.isfinite(te), te, 1e4)
    br, bt, bp = (resample_to_xyz_cube(model[k], x_grid, y_grid, z_grid, phi0_offset, 0.0, r_min) for k in ("br", "bt", "bp"))
    b = np.sqrt(br ** 2 + bt ** 2 + bp ** 2)
    return dict(x_grid=x_grid, y_grid=y_grid, z_grid=z_grid, omega_pe=omega_pe, ne=ne, te=te, b=b, br=br, bt=bt, bp=bp)


def resample_MAS(model, N_pix, X_range, Y_range, N_z, dz0, phi0_offset=0.0, r_min=0.9999999):
    """script/resampling_MAS_LOS.py:141-301 restated (variable z spacing): per-pixel straight LOS
    sampling of rho, te, br, bt, bp; returns the LOS_data dict."""
    R_sun_m, R_sun_cm = 6.957e8, 6.957e10
    idx_z = np.arange(N_z)
    dz = dz0 * (1 + (5 * idx_z / N_z) ** 2.5)
    z_coords = np.cumsum(dz) * R_sun_m
    x_coords = np.linspace(X_range[0], X_range[1], N_pix) * R_sun_m
    y_coords = np.linspace(Y_range[0], Y_range[1], N_pix) * R_sun_m
    X, Y = np.meshgrid(x_coords, y_coords)
    temp = "te" if "te" in model else "t"
    shape = (N_pix, N_pix, N_z)
    Ne, Te, B = (np.full(shape, np.nan) for _ in range(3))
    ds = np.zeros(shape)
    for k in range(N_z):
        ds[:, :, k] = dz[k] * R_sun_cm
    for i in range(N_pix):
        for j in range(N_pix):
            x, y = X[i, j], Y[i, j]
            if np.sqrt(x ** 2 + y ** 2) < R_sun_m:
                z_start = np.sqrt(R_sun_m ** 2 - (x ** 2 + y ** 2)) - 1e-6
            else:
                z_start = -np.sqrt(x ** 2 + y ** 2 - R_sun_m ** 2) - 1e-6
            z_arr = z_start + z_coords
            r_m, colat, lon = cart_to_sph(np.full(N_z, x), -z_arr, np.full(N_z, y), phi0_offset)
            r = r_m / R_sun_m
            valid = r >= r_min
            if not np.any(valid):
                continue
            lat = np.pi / 2 - colat
            Ne[i, j] = sample_at_coords(model["rho"], lon, lat, r)
            Te[i, j] = sample_at_coords(model[temp], lon, lat, r)
            B[i, j] = np.sqrt(sum(sample_at_coords(model[c], lon, lat, r) ** 2 for c in ("br", "bt", "bp")))
            for a in (Ne, Te, B):
                a[i, j, ~valid] = np.nan
    return dict(Ne_LOS=Ne, Te_LOS=Te, B_LOS=B, ds_LOS=ds, x_coords=x_coords, y_coords=y_coords, z_coords=z_coords)
