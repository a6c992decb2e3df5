#!/usr/bin/env python
"""One launch of a stage kernel for ncu: `python scripts/ncu_stage.py c1|c2|c3` (development aid).
c1: sample_paths_kernel on BASELINE config 1; c2: grff_slice_kernel on config 2; c3: trace + sample + emission kernels
on config 3 (staged path)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))     # cases.los_sampler_case: the reference benchmark's fixture
import cases  # noqa: E402
from raytracinggrff_b200 import RaySession, synthetic  # noqa: E402

which = sys.argv[1]
ses = RaySession()
ses.ctx.set_pipeline(False)            # one launch over the whole input
if which == "c1":
    xg, yg, zg, ne, te, b, r_record, s_arr, start = cases.los_sampler_case(256, 256, 128, seed=0)
    ses.set_field_cubes(xg, yg, zg, ne, te, b)
    ses.sample(r_record, s_arr, start, 6.957e10)
    ses.sample(r_record, s_arr, start, 6.957e10)
    print("c1 kernel ms", ses.ctx.last_kernel_ms)
elif which == "c2":
    los = synthetic.straight_los_case(N_pix=256, N_z=400)
    npix, nz, nf = 256 * 256, 400, 4
    ne2, te2, b2, ds2 = (los[k].reshape(npix, nz) for k in ("Ne_LOS", "Te_LOS", "B_LOS", "ds_LOS"))
    valid = ~(np.isnan(ne2) | np.isnan(te2) | np.isnan(b2))
    P = np.zeros((15, nz, npix), order="F")
    P[4], P[6], P[7] = 90.0, 5, 30
    for m, a in ((0, ds2), (1, te2), (2, ne2), (3, b2)):
        P[m] = np.where(valid, a, 0.0).T
    area = (los["x_coords"][1] - los["x_coords"][0]) ** 2 * 1e4
    L = np.array([npix, nz, nf, 1, 0, 0], dtype=np.int32)
    R = np.zeros((3, npix), order="F"); R[0], R[1], R[2] = area, 450e6, 0.1
    RL = np.zeros((7, nf, npix), order="F")
    ses.get_mw_slice(L, R, P, RL)
    ses.get_mw_slice(L, R, P, RL)
    print("c2 kernel ms", ses.ctx.last_kernel_ms)
else:
    c = synthetic.corona_cube(128, 3.0)
    g3 = (c["x_grid"], c["y_grid"], c["z_grid"])
    ses.set_omega_cube(c["omega_pe"], *g3)
    ses.set_field_cubes(*g3, c["ne"], c["te"], c["b"])
    xs, ys, zs, kv = synthetic.ray_launch_geometry(64, 1.44, 3.0)
    for _ in range(2):
        ses.trace(75e6, xs, ys, zs, kv, 6e-3, 5000, 10, True, 2.0, fetch=False)
        ses.sample_traced(np.column_stack([xs, ys, zs]), 6.957e10, fetch=False)
        ses.emission_traced((2 * 1.44 / 64 * 6.957e10) ** 2, 75e6)
    print("c3 emission kernel ms", ses.ctx.last_kernel_ms)
