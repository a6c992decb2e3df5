"""Seeded input cases shared by the golden-vector generator, the oracle tests and the GPU
parity tests.  Inputs are regenerated from code; only reference OUTPUTS are committed under
tests/golden/ (see tests/golden/make_golden.py)."""
from __future__ import annotations

import numpy as np

from raytracinggrff_b200 import synthetic


def sampler_fixture(seed=0):
    """The reference test fixture (/root/reference/tests/test_gpu_raytrace.py:13-44): 33^3 cube on
    [-1,1]^3 with linear/quadratic fields, 128 random straight rays x 64 samples, zeros and NaNs
    planted in S, eight rays forced out of bounds.  Same RNG call order => same arrays."""
    rng = np.random.default_rng(seed)
    n = 33
    g = np.linspace(-1.0, 1.0, n, dtype=np.float32)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    ne = (x + y + z).astype(np.float32)
    te = (x * x + 2.0 * y + 3.0 * z).astype(np.float32)
    b = (2.0 * x - y + 0.5 * z).astype(np.float32)
    n_steps, n_rays = 64, 128
    origin = rng.uniform(-0.8, 0.8, size=(n_rays, 3)).astype(np.float32)
    dirs = rng.normal(size=(n_rays, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    s = (np.arange(n_steps, dtype=np.float32) * 0.03)[:, None]
    r_record = origin[None, :, :] + s[:, :, None] * dirs[None, :, :]
    s_arr = np.ones((n_steps, n_rays), dtype=np.float32)
    s_arr[::9, ::7] = 0.0
    s_arr[::13, ::11] = np.nan
    r_record[-5:, :8, 0] = 2.5
    return g, g.copy(), g.copy(), ne, te, b, r_record, s_arr, origin.copy()


def los_sampler_case(n_pix=256, n_steps=256, grid_n=128, seed=0):
    """BASELINE config 1: the synthetic LOS-sampling case of the reference benchmark
    (bench_raytrace.py:126-151): Gaussian n_e, linear T and B on [-2,2]^3, jittered straight rays
    from z = 2.5, sample spacing 0.02, S = 1.  Same RNG call sequence, hence the same arrays."""
    rng = np.random.default_rng(seed)
    g = np.linspace(-2.0, 2.0, grid_n, dtype=np.float32)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    ne = (1.0e8 + 2.0e8 * np.exp(-(x * x + y * y + z * z))).astype(np.float32)
    te = (1.0e6 + 2.0e6 * (x + 2 * y - z)).astype(np.float32)
    b = (2.0 + x - y + 0.5 * z).astype(np.float32)
    n_rays = n_pix * n_pix
    origin_xy = rng.uniform(-1.2, 1.2, size=(n_rays, 2)).astype(np.float32)
    origin = np.column_stack([origin_xy, np.full(n_rays, 2.5, dtype=np.float32)])
    dirs = np.tile(np.array([[0.0, 0.0, -1.0]], dtype=np.float32), (n_rays, 1))
    dirs[:, 0:2] += rng.normal(scale=0.02, size=(n_rays, 2)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    s = (np.arange(n_steps, dtype=np.float32) * 0.02)[:, None]
    r_record = origin[None, :, :] + s[:, :, None] * dirs[None, :, :]
    s_arr = np.ones((n_steps, n_rays), dtype=np.float32)
    return g, g.copy(), g.copy(), ne, te, b, r_record, s_arr, origin


def trace_case(name):
    """Integrator cases.  Returns kwargs for ray_trace(...)."""
    if name == "corona_cs":
        # config-3 physics on a small cube: 8x8 pixels, cross-sections on, rays reflect and exit
        c = synthetic.corona_cube(48, 3.0)
        xs, ys, zs, kv = synthetic.ray_launch_geometry(8, 1.44, 3.0)
        return dict(omega_pe_3d=c["omega_pe"], x_grid=c["x_grid"], y_grid=c["y_grid"], z_grid=c["z_grid"],
                    freq_hz=75e6, x_start=xs, y_start=ys, z_start=zs, kvec_in_norm=kv, dt=6e-3,
                    n_steps=3000, record_stride=10, trace_crosssections=True, perturb_ratio=2)
    if name == "oblique_nocs":
        # anisotropic grid, oblique float directions, odd stride, starts outside the cube (NaN k),
        # starts exactly on a face, no cross-sections
        rng = np.random.default_rng(7)
        xg = np.linspace(-2.0, 2.5, 37)
        yg = np.linspace(-1.5, 1.5, 29)
        zg = np.linspace(-3.0, 2.0, 41)
        X, Y, Z = np.meshgrid(xg, yg, zg, indexing="ij")
        r = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) + 0.3
        ne = 3e8 * np.exp(-1.5 * (r - 0.3)) * (1 + 0.3 * np.sin(2 * X) * np.cos(Y))
        w = synthetic.omega_pe_from_ne(ne)
        n = 40
        xs = rng.uniform(-1.8, 2.3, n)
        ys = rng.uniform(-1.3, 1.3, n)
        zs = np.full(n, 2.0)
        xs[:3] = [2.6, -2.1, 0.0]          # two starts outside in x, one fine
        zs[3] = 2.0000001                  # just outside the top face
        kv = np.column_stack([rng.normal(scale=0.2, size=n), rng.normal(scale=0.2, size=n), -np.ones(n)])
        kv /= np.linalg.norm(kv, axis=1, keepdims=True)
        return dict(omega_pe_3d=w, x_grid=xg, y_grid=yg, z_grid=zg, freq_hz=120e6, x_start=xs,
                    y_start=ys, z_start=zs, kvec_in_norm=kv, dt=4e-3, n_steps=1500, record_stride=7,
                    trace_crosssections=False, perturb_ratio=2)
    if name == "c3_subset":
        # BASELINE config 3 exactly (128^3, extent 3, 64^2 image, 75 MHz, dt 6e-3, 5000 steps,
        # stride 10, cross-sections, perturb 2) restricted to every 8th pixel in x and y
        c = synthetic.corona_cube(128, 3.0)
        xs, ys, zs, kv = synthetic.ray_launch_geometry(64, 1.44, 3.0)
        sel = (np.arange(64)[::8][:, None] * 64 + np.arange(64)[::8][None, :]).ravel()
        return dict(omega_pe_3d=c["omega_pe"], x_grid=c["x_grid"], y_grid=c["y_grid"], z_grid=c["z_grid"],
                    freq_hz=75e6, x_start=xs[sel], y_start=ys[sel], z_start=zs[sel], kvec_in_norm=kv[sel],
                    dt=6e-3, n_steps=5000, record_stride=10, trace_crosssections=True, perturb_ratio=2)
    raise KeyError(name)


TRACE_CASES = ("corona_cs", "oblique_nocs", "c3_subset")


def image_case(name):
    """Emission maps with failed pixels for the image-plane operations (T_b-like values)."""
    rng = np.random.default_rng({"sparse": 11, "clusters": 12, "cube": 13, "rows": 14}[name])
    if name == "sparse":
        a = 1e6 * (1.0 + rng.random((37, 45)))
        a[rng.random(a.shape) < 0.03] = np.nan
        a[5, 7] = np.inf
        return a
    if name == "clusters":
        # runs of NaN along rows and columns, a NaN block, NaN corners and a NaN border row
        a = 5e5 + 1e5 * rng.standard_normal((40, 33))
        a[10, 3:20] = np.nan
        a[4:30, 25] = np.nan
        a[20:26, 8:15] = np.nan
        a[0, :] = np.nan
        a[-1, -1] = np.nan
        a[0, 0] = np.nan
        return a
    if name == "rows":
        # whole rows and columns missing: some pixels have no finite pixel in one or two directions
        a = 1e6 * rng.random((21, 19))
        a[:, 4] = np.nan
        a[7, :] = np.nan
        a[:3, :3] = np.nan
        return a
    if name == "cube":
        # (ny, nx, nf): one clean plane, one sparse, one with nothing finite at all
        a = 1e6 * (1.0 + rng.random((24, 28, 3)))
        m = rng.random((24, 28)) < 0.1
        a[m, 1] = np.nan
        a[:, :, 2] = np.nan
        return a
    raise KeyError(name)


IMAGE_CASES = ("sparse", "clusters", "rows", "cube")
BEAM_SIGMAS = (0.0, 0.7, 2.5, 11.0, 60.0)
