#!/usr/bin/env python
"""Per-stage timings of BASELINE configs 1-3 against their rooflines with the CPU oracle beside
them.  bench.py appends `run_all(...)` to its JSON line (extras.stages_configs_1_to_3), so every BASELINE
config has a driver-run figure; run directly it prints the same dict.

  C1  LOS sampler      65 536 rays x 256 samples, 128^3 cube        (bench_raytrace.py shape)
  C2  GRFF batched     256^2 pixels x 400 z samples x 4 freqs        (straight-LOS shape)
  C3  ray-traced map   64^2 pixels, 128^3 cube, 5000 steps, stride 10 (staged: trace -> sample -> emission, and fused)

"kernel_ms" are CUDA-event times of the kernel alone (rtgrff_ctx_last_kernel_ms with the chunked host pipeline
switched off, so the events bracket nothing but the kernel); "e2e_ms" is the wall time of the public API call with
pageable host (numpy) arrays in and out, chunked pipeline on (the default).
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))     # cases.los_sampler_case: the reference benchmark's fixture
import cases  # noqa: E402
from raytracinggrff_b200 import RaySession, synthetic  # noqa: E402

HBM = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def best(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def run_all(ses, quick=False, with_cube_builder=True):
    from oracle import oracle
    oracle.build()
    threads = oracle.set_num_threads(None)
    reps = 2 if quick else 5
    out = {"hbm_peak_gbs": HBM, "cpu_threads": threads}

    def kernel_only(fn):
        ses.ctx.set_pipeline(False)
        try:
            fn(); fn()
            return ses.ctx.last_kernel_ms * 1e-3
        finally:
            ses.ctx.set_pipeline(True)

    # ---- C1 sampler -------------------------------------------------------------------------
    args = cases.los_sampler_case(256, 256, 128, seed=0)
    xg, yg, zg, ne, te, b, r_record, s_arr, start = args
    ses.set_field_cubes(xg, yg, zg, ne, te, b)
    n = r_record.shape[0] * r_record.shape[1]
    run1 = lambda: ses.sample(r_record, s_arr, start, 6.957e10)   # noqa: E731
    k = kernel_only(run1)
    e2e = best(run1, reps)
    t0 = time.perf_counter(); oracle.sample_model_with_rays_cpu(*args, r_sun_cm=6.957e10); t_cpu = time.perf_counter() - t0
    out["C1_sampler"] = {"samples": n, "kernel_ms": k * 1e3, "samples_per_s_kernel": n / k,
                         "alg_GBps": n * 129 / k / 1e9, "frac_hbm": n * 129 / k / 1e9 / HBM,
                         "e2e_ms": e2e * 1e3, "samples_per_s_e2e": n / e2e, "e2e_host_bytes": int(n * 33),
                         "e2e_host_GBps": n * 33 / e2e / 1e9, "cpu_oracle_s": t_cpu, "cpu_samples_per_s": n / t_cpu,
                         "reference_numpy_single_core_samples_per_s": 3.55e6}
    del args, r_record, s_arr

    # ---- C2 GRFF batched --------------------------------------------------------------------
    los = synthetic.straight_los_case(N_pix=256, N_z=400)
    npix, nz, nf = 256 * 256, 400, 4
    ne2, te2, b2, ds2 = (los[kk].reshape(npix, nz) for kk in ("Ne_LOS", "Te_LOS", "B_LOS", "ds_LOS"))
    valid = ~(np.isnan(ne2) | np.isnan(te2) | np.isnan(b2))
    P = np.zeros((15, nz, npix), order="F")
    P[4], P[6], P[7] = 90.0, 5, 30
    for m, a in ((0, ds2), (1, te2), (2, ne2), (3, b2)):
        P[m] = np.where(valid, a, 0.0).T
    area = (los["x_coords"][1] - los["x_coords"][0]) ** 2 * 1e4
    L = np.array([npix, nz, nf, 1, 0, 0], dtype=np.int32)
    R = np.zeros((3, npix), order="F"); R[0], R[1], R[2] = area, 450e6, 0.1
    RL = np.zeros((7, nf, npix), order="F")
    run2 = lambda: ses.get_mw_slice(L, R, P, RL)    # noqa: E731
    k = kernel_only(run2)
    e2e = best(run2, 2 if quick else 3)
    RL_ref = np.zeros_like(RL)
    t0 = time.perf_counter(); oracle.get_mw_slice(L, R, P, None, None, None, RL_ref); t_cpu = time.perf_counter() - t0
    ev = npix * nz * nf
    err = float(np.max(np.abs(RL[5:] - RL_ref[5:]) / (np.abs(RL_ref[5:]).max())))
    out["C2_grff_slice"] = {"voxel_freq_evals": ev, "kernel_ms": k * 1e3, "evals_per_s_kernel": ev / k,
                            "parms_GBps": P.nbytes / k / 1e9, "frac_hbm": P.nbytes / k / 1e9 / HBM,
                            "e2e_ms": e2e * 1e3, "e2e_host_bytes": int(P.nbytes + RL.nbytes), "e2e_host_GBps": (P.nbytes + RL.nbytes) / e2e / 1e9,
                            "cpu_oracle_s": t_cpu, "cpu_evals_per_s": ev / t_cpu, "max_rel_err_vs_oracle": err}
    del P, RL, RL_ref, los

    # ---- C3 staged and fused ----------------------------------------------------------------
    c = synthetic.corona_cube(128, 3.0)
    g3 = (c["x_grid"], c["y_grid"], c["z_grid"])
    ses.set_omega_cube(c["omega_pe"], *g3)
    ses.set_field_cubes(*g3, c["ne"], c["te"], c["b"])
    xs, ys, zs, kv = synthetic.ray_launch_geometry(64, 1.44, 3.0)
    start = np.column_stack([xs, ys, zs])
    area = (2 * 1.44 / 64 * 6.957e10) ** 2
    st = {}

    def staged():
        _, _, act = ses.trace(75e6, xs, ys, zs, kv, 6e-3, 5000, 10, True, 2.0, fetch=False); st["trace_ms"] = ses.ctx.last_kernel_ms; st["act"] = act
        ses.sample_traced(start, 6.957e10, fetch=False); st["sample_ms"] = ses.ctx.last_kernel_ms
        tb, vi = ses.emission_traced(area, 75e6); st["emission_ms"] = ses.ctx.last_kernel_ms
        st["tb_staged"] = tb[:, 0]
    e2e_staged = best(staged, 3)

    def fused():
        tb, _, s2 = ses.render_map(xs, ys, zs, [(75e6, 6e-3, 5000, 10)], kvec_in_norm=kv, pixel_area_cm2=area); st["fused_ms"] = ses.ctx.last_kernel_ms
        st["tb_fused"] = tb[0]
    e2e_fused = best(fused, 3)
    t0 = time.perf_counter()
    r, cs = oracle.ray_trace(c["omega_pe"], *g3, 75e6, xs, ys, zs, kv, 6e-3, 5000, 10, True, perturb_ratio=2)
    t_tr = time.perf_counter() - t0
    smp = oracle.sample_model_with_rays_cpu(*g3, c["ne"], c["te"], c["b"], r, np.array(cs), start, 6.957e10)
    t_sm = time.perf_counter() - t0 - t_tr
    tb_ref, _, _ = oracle.emission_from_samples(smp, 64, 1.44, 75e6)
    t_em = time.perf_counter() - t0 - t_tr - t_sm
    tb_ref = tb_ref.ravel()
    nzm = tb_ref != 0
    nominal = 4096 * 5000
    out["C3_map"] = {"nominal_ray_steps": nominal, "active_ray_steps": st["act"],
                     **{k2: v for k2, v in st.items() if k2 in ("trace_ms", "sample_ms", "emission_ms", "fused_ms")},
                     "e2e_staged_ms": e2e_staged * 1e3, "e2e_fused_ms": e2e_fused * 1e3,
                     "nominal_ray_steps_per_s_trace_kernel": nominal / (st["trace_ms"] * 1e-3),
                     "nominal_ray_steps_per_s_fused_e2e": nominal / e2e_fused,
                     "cpu_oracle_trace_s": t_tr, "cpu_oracle_sample_s": t_sm, "cpu_oracle_emission_s": t_em,
                     "cpu_nominal_ray_steps_per_s": nominal / t_tr,
                     "max_rel_dTb_fused_vs_oracle": float(np.max(np.abs(st["tb_fused"][nzm] - tb_ref[nzm]) / tb_ref[nzm])),
                     "max_rel_dTb_staged_vs_oracle": float(np.max(np.abs(st["tb_staged"][nzm] - tb_ref[nzm]) / tb_ref[nzm])),
                     "reference_numpy_single_core_ray_steps_per_s": 2.02e5}
    if with_cube_builder and not quick:
        # ---- cube builder: spherical model -> 256^3 cubes ----------------------
        from oracle import oracle_cubes as oc
        m = synthetic.spherical_corona(150, 110, 128, active_region=True)
        g = np.linspace(-3.0, 3.0, 256)
        t_gpu = best(lambda: ses.set_model_from_spherical(m, g, g, g, phi0_offset=24.0, want_bvec=True), n=3)
        g_small = np.linspace(-3.0, 3.0, 64)
        t0 = time.perf_counter(); oc.compose_cubes(m, g_small, g_small, g_small, phi0_offset=24.0); t_cpu64 = time.perf_counter() - t0
        out["cube_builder"] = {"cube": "256^3 x {rho,te,br,bt,bp} from a 128x110x150 (phi,lat,r) mesh", "gpu_e2e_ms": t_gpu * 1e3,
                               "voxel_vars_per_s": 5 * 256 ** 3 / t_gpu, "cpu_oracle_64cube_s": t_cpu64,
                               "cpu_oracle_voxel_vars_per_s": 5 * 64 ** 3 / t_cpu64,
                               "note": "oracle = numpy/scipy restatement (single thread); the reference's psipy loop is minutes per cube"}
    return out


def main():
    ses = RaySession()
    print(json.dumps(run_all(ses, quick="--quick" in sys.argv), indent=1))


if __name__ == "__main__":
    main()
