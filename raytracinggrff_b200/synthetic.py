"""Deterministic synthetic inputs for the per-ray hot path (host-side numpy only).

The reference builds its cubes from MAS model files through psipy
(script/resample_with_ray_tracing.py:251-293), which is out of scope.
What is kept from there are the *formulas* that turn resampled fields into the arrays the hot
path consumes:

* ``omega_pe = 2*pi*8.93e3*sqrt(max(n_e, 0))``, NaN -> 0     (script/...:271-273)
* ``n_e >= 0``; ``T`` NaN -> 1e4; ``|B| = sqrt(br^2+bt^2+bp^2)``  (script/...:279-293)
* r < R_MIN gets n_e = 0, B = 0, T = 1e4 through the NaN fills   (script/...:71, :124, :150)

and the ray launch geometry (script/...:295-303), the frequency grid (:355) and the per-frequency
integration presets of the publication drivers (script/pub/compare_LOS_raytracing.py:35-63).
The analytic corona itself is this repo's: Newkirk-like density with an
equatorial streamer enhancement, a tanh temperature profile and a dipole (+ optional compact
active-region dipole so that gyroresonance matters at GHz frequencies).
"""
from __future__ import annotations

import functools

import numpy as np

R_MIN = 0.999999          # script/resample_with_ray_tracing.py:71
R_SUN_CM = 6.957e10       # script/resample_with_ray_tracing.py:68
R_SUN_M = 6.957e8         # script/resample_with_ray_tracing.py:69


def omega_pe_from_ne(ne):
    """script/resample_with_ray_tracing.py:271-273."""
    w = 8.93e3 * np.sqrt(np.maximum(ne, 0.0)) * 2 * np.pi
    return np.nan_to_num(w, nan=0.0, posinf=0.0, neginf=0.0)


def corona_cube(grid_n, extent, active_region=False, b0=2.0, dtype=np.float64, x_slice=None):
    """Analytic corona on ``linspace(-extent, extent, grid_n)^3`` (x slowest, z fastest; the
    observer sits on +z and solar north is +y).  Returns a dict with the 1-D grids and the cubes
    ``ne, te, b, bx, by, bz, omega_pe``.  ``x_slice``: only those x planes of the cubes (a 512^3 cube in
    slabs keeps the temporaries small); the grids stay the full ones."""
    g = np.linspace(-extent, extent, grid_n)
    x = (g if x_slice is None else g[x_slice])[:, None, None]
    y = g[None, :, None]
    z = g[None, None, :]
    r2 = x * x + y * y + z * z
    r = np.sqrt(r2)
    inside = r < R_MIN
    rs = np.where(inside, 1.0, r)                      # avoid 1/0 where the value is discarded
    lat = np.degrees(np.arcsin(np.clip(y / rs, -1.0, 1.0)))
    ne = 4.2e4 * 10.0 ** (4.32 / rs) * (1.0 + 0.5 * np.exp(-(lat / 15.0) ** 2))
    ne = np.where(inside, 0.0, ne)
    te = np.where(inside, 1.0e4, 1.0e6 + 0.4e6 * np.tanh(rs - 1.0))
    # dipole along +y:  B = b0/r^3 [3 (m.rhat) rhat - m]
    mr = y / rs
    f = b0 / rs ** 3
    bx = f * (3.0 * mr * x / rs)
    by = f * (3.0 * mr * y / rs - 1.0)
    bz = f * (3.0 * mr * z / rs)
    if active_region:
        # compact dipole buried 0.05 R_sun below the surface on the observer-facing side
        c = np.array([0.30, 0.20, np.sqrt(1.0 - 0.30 ** 2 - 0.20 ** 2)]) * 0.95
        m = c / np.linalg.norm(c)
        d = 0.05
        px, py, pz = x - c[0], y - c[1], z - c[2]
        pr = np.sqrt(px * px + py * py + pz * pz)
        pr = np.maximum(pr, 0.25 * d)
        md = (m[0] * px + m[1] * py + m[2] * pz) / pr
        fa = 300.0 * (d / pr) ** 3
        bx = bx + fa * (3.0 * md * px / pr - m[0])
        by = by + fa * (3.0 * md * py / pr - m[1])
        bz = bz + fa * (3.0 * md * pz / pr - m[2])
    bx = np.where(inside, 0.0, bx)
    by = np.where(inside, 0.0, by)
    bz = np.where(inside, 0.0, bz)
    b = np.sqrt(bx * bx + by * by + bz * bz)
    out = dict(x_grid=g, y_grid=g.copy(), z_grid=g.copy(), ne=ne, te=te, b=b, bx=bx, by=by, bz=bz,
               omega_pe=omega_pe_from_ne(ne))
    if dtype != np.float64:
        for k in ("ne", "te", "b", "bx", "by", "bz"):
            out[k] = out[k].astype(dtype)
    return out


def ray_launch_geometry(N_pix, X_fov, z_observer, N_pix_y=None):
    """Image grid and ray starts, script/resample_with_ray_tracing.py:295-303.
    Ray p = i*N_pix + j starts at (x[j], y[i], z_start) with direction (0,0,-1)."""
    N_pix_y = N_pix if N_pix_y is None else N_pix_y
    x_coords = np.linspace(-X_fov, X_fov, N_pix)
    y_coords = np.linspace(-X_fov, X_fov, N_pix_y)
    X_img, Y_img = np.meshgrid(x_coords, y_coords)
    x_flat = X_img.ravel()
    y_flat = Y_img.ravel()
    z_start = np.sqrt(np.abs((z_observer * 2.0) ** 2 - x_flat ** 2 - y_flat ** 2)) / 2.0
    kvec_in_norm = np.tile([[0, 0, -1]], (len(x_flat), 1))
    return x_flat, y_flat, z_start, kvec_in_norm


def log_frequencies(freq0, n_freq, log_step):
    """script/resample_with_ray_tracing.py:355."""
    return freq0 * (10.0 ** (log_step * np.arange(n_freq)))


def frequency_scaled_params(freq_hz, ref_freq_hz=100e6, base_dt=6e-3, base_n_steps=4000,
                            base_record_stride=5, scaling_exp=0.5, min_n_steps=1200):
    """Per-frequency integrator settings of the publication drivers
    (script/pub/compare_LOS_raytracing.py:35-63, script/pub/TbSpectra_gen.py:27-44)."""
    scale = (ref_freq_hz / freq_hz) ** scaling_exp
    return {
        "dt": base_dt * scale,
        "n_steps": max(min_n_steps, int(round(base_n_steps / max(scale, 1e-12)))),
        "record_stride": max(1, int(round(base_record_stride * scale))),
    }


def straight_los_case(N_pix=256, N_z=400, X_fov=1.44, dz0=3e-4, r_min=0.9999999, b0=2.0):
    """BASELINE config 2: the LOS_data.npz arrays of script/resampling_MAS_LOS.py filled from the
    analytic corona.  Irregular z grid dz = dz0 (1+(5 i/N_z)^2.5) (:141-146); every LOS starts at
    the solar surface / plane of sky (:198-201) and runs toward the observer; samples with
    r < r_min are NaN (:210-218); ds = dz R_sun_cm (:187-188).  Sample order is Sun -> observer."""
    idx = np.arange(N_z)
    dz = dz0 * (1 + (5 * idx / N_z) ** 2.5)
    zc = np.cumsum(dz)
    xs = np.linspace(-X_fov, X_fov, N_pix)
    X, Y = np.meshgrid(xs, xs)
    rho2 = X ** 2 + Y ** 2
    z_start = np.where(rho2 < 1.0, np.sqrt(np.abs(1.0 - rho2)), -np.sqrt(np.abs(rho2 - 1.0))) - 1e-6 / R_SUN_M
    Z = z_start[:, :, None] + zc[None, None, :]
    r = np.sqrt(rho2[:, :, None] + Z ** 2)
    ok = r >= r_min
    rs = np.where(ok, r, 1.0)
    lat = np.degrees(np.arcsin(np.clip(Y[:, :, None] / rs, -1.0, 1.0)))
    ne = 4.2e4 * 10.0 ** (4.32 / rs) * (1.0 + 0.5 * np.exp(-(lat / 15.0) ** 2))
    te = 1.0e6 + 0.4e6 * np.tanh(rs - 1.0)
    mr = Y[:, :, None] / rs
    b = b0 / rs ** 3 * np.sqrt(1.0 + 3.0 * mr ** 2)
    nan = np.nan
    out = dict(Ne_LOS=np.where(ok, ne, nan), Te_LOS=np.where(ok, te, nan), B_LOS=np.where(ok, b, nan),
               ds_LOS=np.broadcast_to(dz * R_SUN_CM, r.shape).copy(),
               x_coords=xs * R_SUN_M, y_coords=xs * R_SUN_M, z_coords=zc * R_SUN_M)
    return out


@functools.lru_cache(maxsize=16)
def tile_order(n_x, n_y, tile_w=8, tile_h=4):
    """Permutation of the flat pixel indices p = i*n_x + j that walks the image in tile_w x tile_h
    pixel tiles (row-major inside a tile): with one thread per ray, a warp of 32 consecutive rays
    then covers an 8x4 pixel patch instead of a 32x1 strip, so its gathers touch fewer cube cells.
    Returns `perm` such that rays[perm] is the tiled order; scatter results back with
    ``out[perm] = out_tiled``."""
    i = np.arange(n_y)[:, None]
    j = np.arange(n_x)[None, :]
    tiles_x = (n_x + tile_w - 1) // tile_w
    key = ((i // tile_h) * tiles_x + (j // tile_w)) * (tile_w * tile_h) + (i % tile_h) * tile_w + (j % tile_w)
    return np.argsort(key.ravel(), kind="stable")


def spherical_corona(n_r=80, n_lat=60, n_phi=90, r_max=6.0, active_region=False, b0=2.0, stagger=True):
    """The analytic corona of `corona_cube` sampled on a MAS-like spherical mesh: non-uniform in r
    (geometric stretch away from the surface) and in latitude (finer near the equator), uniform in
    longitude, with br / bt / bp on half-cell staggered meshes like MAS.  Values are stored in
    physical units (scale 1).  Model frame as in build_rays.py:93: cube (x, y, z) = model
    (x, z, -y); the solar axis is the model's z."""
    from .cubes import SphericalVariable

    def mesh(shift_r=False, shift_t=False, shift_p=False):
        xi = np.linspace(0.0, 1.0, n_r)
        r = 1.0 + (r_max - 1.0) * (np.expm1(3.0 * xi) / np.expm1(3.0))
        r[0] = 0.9995
        if shift_r:
            r = np.concatenate([[0.999], 0.5 * (r[1:] + r[:-1]), [r_max * 1.001]])
        eta = np.linspace(-1.0, 1.0, n_lat)
        lat = 0.5 * np.pi * (0.7 * eta + 0.3 * eta ** 3)
        if shift_t:
            lat = np.concatenate([[lat[0]], 0.5 * (lat[1:] + lat[:-1]), [lat[-1]]])
        phi = np.linspace(0.0, 2 * np.pi, n_phi, endpoint=False)
        if shift_p:
            phi = phi + 0.5 * (phi[1] - phi[0])
        return phi, lat, r

    def fields(phi, lat, r):
        P, T, R = np.meshgrid(phi, lat, r, indexing="ij")
        zc = R * np.sin(T)                      # model z (solar axis)
        xm = R * np.cos(T) * np.cos(P)
        ym = R * np.cos(T) * np.sin(P)
        cx, cy, cz = xm, zc, -ym                # cube frame
        rs = np.maximum(R, 1e-3)
        latd = np.degrees(np.arcsin(np.clip(cy / rs, -1.0, 1.0)))
        ne = 4.2e4 * 10.0 ** (4.32 / rs) * (1.0 + 0.5 * np.exp(-(latd / 15.0) ** 2))
        te = 1.0e6 + 0.4e6 * np.tanh(rs - 1.0)
        mr = cy / rs
        f = b0 / rs ** 3
        bx, by, bz = f * (3 * mr * cx / rs), f * (3 * mr * cy / rs - 1.0), f * (3 * mr * cz / rs)
        if active_region:
            c = np.array([0.30, 0.20, np.sqrt(1.0 - 0.30 ** 2 - 0.20 ** 2)]) * 0.95
            m = c / np.linalg.norm(c)
            d = 0.05
            px, py, pz = cx - c[0], cy - c[1], cz - c[2]
            pr = np.maximum(np.sqrt(px * px + py * py + pz * pz), 0.25 * d)
            md = (m[0] * px + m[1] * py + m[2] * pz) / pr
            fa = 300.0 * (d / pr) ** 3
            bx = bx + fa * (3 * md * px / pr - m[0]); by = by + fa * (3 * md * py / pr - m[1]); bz = bz + fa * (3 * md * pz / pr - m[2])
        # cube-frame vector -> model frame -> (br, bt, bp) with theta the colatitude from the model z
        vx, vy, vz = bx, -bz, by
        st, ct = np.cos(T), np.sin(T)           # colatitude: sin(theta) = cos(lat)
        cp, sp = np.cos(P), np.sin(P)
        br = vx * st * cp + vy * st * sp + vz * ct
        bt = vx * ct * cp + vy * ct * sp - vz * st
        bp = -vx * sp + vy * cp
        return ne, te, br, bt, bp

    out = {}
    phi, lat, r = mesh()
    ne, te, _, _, _ = fields(phi, lat, r)
    out["rho"] = SphericalVariable(ne.astype(np.float32), phi, lat, r)
    out["te"] = SphericalVariable(te.astype(np.float32), phi, lat, r)
    for name, idx, sh in (("br", 2, dict(shift_r=stagger)), ("bt", 3, dict(shift_t=stagger)), ("bp", 4, dict(shift_p=stagger))):
        phi, lat, r = mesh(**sh)
        out[name] = SphericalVariable(fields(phi, lat, r)[idx].astype(np.float32), phi, lat, r)
    return out
