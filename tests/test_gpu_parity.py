"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through the C ABI
of librtgrff_b200.so via the drop-in Python API and is compared with the CPU oracle on the same
seeded inputs and with the golden vectors produced by the unmodified reference.

Tolerances (BASELINE.json north_star):
  ray endpoints / path samples   <= 1e-5 R_sun absolute
  cross-section ratio S          <= 1e-4 relative
  sampler ne/te/b                bit-exact (reference GPU-vs-CPU tolerance is 1e-5), ds 1e-6 relative
  T_b, Stokes I, V               <= 1e-4 relative
"""
import numpy as np
import pytest

import cases
import grff_checks
from raytracinggrff_b200 import synthetic

pytestmark = pytest.mark.gpu

POS_TOL = 1e-5
S_RTOL = 1e-4
TB_RTOL = 1e-4


def _cmp_paths(r, s, r_ref, s_ref, lo=None, hi=None, step_len=None):
    """Paths against the oracle.  Records taken while the oracle's ray is inside the cube must agree
    to POS_TOL.  The step on which a ray LEAVES the cube is discontinuous in the reference itself: an
    RK4 stage that lands a rounding error inside or outside the face keeps or loses its whole
    derivative (build_rays.py:169-174), which moves the frozen exit position by up to a step length.
    With `lo/hi/step_len` given, frozen post-exit records may therefore differ by up to one step
    length, on at most 0.2 % of the rays; without them every record is held to POS_TOL."""
    assert r.shape == r_ref.shape and r.dtype == np.float64
    nan_a, nan_b = np.isnan(r), np.isnan(r_ref)
    assert np.array_equal(nan_a, nan_b)
    d = np.abs(np.nan_to_num(r) - np.nan_to_num(r_ref)).max(axis=2)
    if lo is None:
        assert d.max() <= POS_TOL, f"max |dr| = {d.max():.3e} R_sun"
        s_ok = np.isfinite(s_ref) if s_ref is not None else None
    else:
        inside = np.all((r_ref >= lo) & (r_ref <= hi), axis=2)
        assert d[inside].max() <= POS_TOL, f"max |dr| inside the cube = {d[inside].max():.3e} R_sun"
        bad_rays = np.flatnonzero((d > POS_TOL).any(axis=0))
        assert d.max() <= step_len, f"max |dr| after exit = {d.max():.3e} R_sun"
        assert bad_rays.size <= max(1, int(2e-3 * r.shape[1])), f"{bad_rays.size} rays differ after their exit"
        s_ok = np.isfinite(s_ref) & inside if s_ref is not None else None
    if s_ref is not None:
        assert np.array_equal(np.isnan(s), np.isnan(s_ref))
        if s_ok.any():
            rel = np.abs(s[s_ok] - s_ref[s_ok]) / np.abs(s_ref[s_ok])
            assert rel.max() <= S_RTOL, f"max rel dS = {rel.max():.3e}"
    return d.max()


def _cube_bounds(kw):
    lo = np.array([kw["x_grid"][0], kw["y_grid"][0], kw["z_grid"][0]])
    hi = np.array([kw["x_grid"][-1], kw["y_grid"][-1], kw["z_grid"][-1]])
    return dict(lo=lo, hi=hi, step_len=1.01 * (2.998e10 / 6.96e10) * kw["dt"])


# ------------------------------------------------------------------------------ integrator ----
@pytest.mark.parametrize("name", cases.TRACE_CASES)
def test_trace_ray_matches_reference_golden(golden, name):
    from raytracinggrff_b200 import trace_ray
    kw = cases.trace_case(name)
    g = golden(f"trace_{name}")
    r, cs = trace_ray("cuda", **kw)
    s_ref = g["s_record"] if kw["trace_crosssections"] else None
    s = np.array(cs) if kw["trace_crosssections"] else None
    if not kw["trace_crosssections"]:
        assert cs == []
    _cmp_paths(r, s, g["r_record"], s_ref, **_cube_bounds(kw))


def test_ray_trace_full_config3_matches_oracle(oracle):
    """BASELINE config 3 at full size (64^2 rays, 128^3 cube, 5000 steps) against the oracle."""
    from raytracinggrff_b200 import ray_trace
    c = synthetic.corona_cube(128, 3.0)
    xs, ys, zs, kv = synthetic.ray_launch_geometry(64, 1.44, 3.0)
    kw = dict(omega_pe_3d=c["omega_pe"], x_grid=c["x_grid"], y_grid=c["y_grid"], z_grid=c["z_grid"], freq_hz=75e6,
              x_start=xs, y_start=ys, z_start=zs, kvec_in_norm=kv, dt=6e-3, n_steps=5000, record_stride=10,
              trace_crosssections=True, perturb_ratio=2)
    r_ref, cs_ref = oracle.ray_trace(**kw)
    r, cs = ray_trace(**kw)
    assert isinstance(cs, list) and len(cs) == 500 and cs[0].shape == (4096,)
    _cmp_paths(r, np.array(cs), r_ref, np.array(cs_ref), **_cube_bounds(kw))


def test_trace_edge_cases(oracle, session):
    g = np.linspace(-1.0, 1.0, 9)
    w = np.full((9, 9, 9), 2e8)
    session.set_omega_cube(w, g, g, g)
    # zero rays / zero steps
    r, s, act = session.trace(80e6, np.zeros(0), np.zeros(0), np.zeros(0), np.zeros((0, 3)), 1e-2, 10, 3, True)
    assert r.shape == (4, 0, 3) and s.shape == (4, 0) and act == 0
    r, s, act = session.trace(80e6, np.zeros(2), np.zeros(2), np.zeros(2), None, 1e-2, 0, 3, False)
    assert r.shape == (0, 2, 3) and s is None
    # evanescent start (omega < omega_pe => kc0 = 0 => omega > 0 but k = 0: ray is pushed by grad only = 0 here)
    xs = np.array([0.0, 0.3]); ys = np.array([0.0, -0.2]); zs = np.array([0.5, 0.9])
    kv = np.array([[0, 0, -1.0], [0, 0, -1.0]])
    r, s, _ = session.trace(10e6, xs, ys, zs, kv, 1e-2, 20, 1, True)
    r_ref, cs_ref = oracle.ray_trace(w, g, g, g, 10e6, xs, ys, zs, kv, 1e-2, 20, 1, True)
    _cmp_paths(r, s, r_ref, np.array(cs_ref))
    # NaN start position and start outside the cube: NaN k, frozen, S NaN
    xs = np.array([np.nan, 1.5, 0.0])
    r, s, act = session.trace(300e6, xs, np.zeros(3), np.zeros(3), np.tile([[0, 0, -1.0]], (3, 1)), 1e-2, 30, 7, True)
    r_ref, cs_ref = oracle.ray_trace(w, g, g, g, 300e6, xs, np.zeros(3), np.zeros(3), np.tile([[0, 0, -1.0]], (3, 1)),
                                     1e-2, 30, 7, True)
    _cmp_paths(r, s, r_ref, np.array(cs_ref))
    assert np.all(np.isnan(s[:, :2]))
    # cumulative S = running product of the per-step ratios
    c = synthetic.corona_cube(32, 3.0)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    xs, ys, zs, kv = synthetic.ray_launch_geometry(4, 1.0, 3.0)
    _, s_step, _ = session.trace(75e6, xs, ys, zs, kv, 6e-3, 200, 1, True, 2.0, s_mode=0)
    _, s_cum, _ = session.trace(75e6, xs, ys, zs, kv, 6e-3, 200, 1, True, 2.0, s_mode=1)
    np.testing.assert_allclose(s_cum, np.cumprod(s_step, axis=0), rtol=1e-12)


def test_trace_invariants_at_scale(session):
    """Size-independent properties on a config-4-sized launch (512^2 rays, 256^3 cube):
    vacuum rays are straight with |dr/dt| = C_R, S = 1; the image-plane mirror symmetry of the
    corona (x -> -x) maps ray p to its mirror ray."""
    from raytracinggrff_b200 import C_R
    n = 256
    g = np.linspace(-3.0, 3.0, n)
    session.set_omega_cube(np.zeros((n, n, n)), g, g, g)
    xs, ys, zs, kv = synthetic.ray_launch_geometry(512, 1.44, 3.0)
    dt, steps, stride = 6e-3, 600, 100
    r, s, act = session.trace(75e6, xs, ys, zs, kv, dt, steps, stride, True, 2.0)
    t = dt * (np.arange(0, steps, stride) + 1)
    expect_z = zs[None, :] - C_R * t[:, None]
    assert np.abs(r[..., 2] - expect_z).max() < 1e-11
    assert np.abs(r[..., 0] - xs[None]).max() == 0 and np.abs(r[..., 1] - ys[None]).max() == 0
    # the pencil basis is normalised with MUFU.RSQ (2 ulp): S = 1 to a few 1e-7 (tolerance on S: 1e-4)
    assert np.abs(s - 1.0).max() < 2e-6
    assert act == steps * xs.size
    c = synthetic.corona_cube(n, 3.0)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    r, s, act = session.trace(75e6, xs, ys, zs, kv, dt, 3000, 300, False)
    img = r.reshape(r.shape[0], 512, 512, 3)
    mirror = img[:, :, ::-1, :] * np.array([-1.0, 1.0, 1.0])
    inside = np.all(np.abs(img) <= 3.0, axis=3) & np.all(np.abs(mirror) <= 3.0, axis=3)
    assert inside.mean() > 0.3
    assert np.abs(img - mirror)[inside].max() < 1e-6
    # frozen exit positions: the exit step is discontinuous (see _cmp_paths), one step length at most
    assert np.abs(img - mirror).max() < 1.01 * C_R * dt
    assert 0 < act < 3000 * xs.size


# --------------------------------------------------------------------------------- sampler ----
@pytest.mark.parametrize("seed", (1, 2, 3))
def test_sampler_matches_reference_fixture(oracle, golden, seed):
    """The reference's GPU-vs-CPU test (tests/test_gpu_raytrace.py:91-110) with its tolerances, plus
    bit-exactness against the reference CPU output."""
    from raytracinggrff_b200 import sample_model_with_rays
    args = cases.sampler_fixture(seed)
    gpu = sample_model_with_rays("cuda", *args, r_sun_cm=1.0)
    cpu = golden(f"sampler_fixture_seed{seed}")
    assert np.array_equal(cpu["valid_mask"], gpu["valid_mask"])
    assert gpu["valid_mask"].dtype == bool
    np.testing.assert_allclose(cpu["ne"], gpu["ne"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cpu["te"], gpu["te"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cpu["b"], gpu["b"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cpu["ds"], gpu["ds"], rtol=1e-6, atol=1e-6)
    for k in ("ne", "te", "b", "ds", "s"):
        assert gpu[k].dtype == np.float32
        assert np.array_equal(cpu[k], gpu[k], equal_nan=True), k
    orc = oracle.sample_model_with_rays_cpu(*args, r_sun_cm=1.0)
    for k in ("ne", "te", "b", "ds", "valid_mask"):
        assert np.array_equal(orc[k], gpu[k], equal_nan=True), k


def test_sampler_config1_full_size(oracle):
    """BASELINE config 1 at full size: 65 536 rays x 256 samples, 128^3 cube."""
    from raytracinggrff_b200 import sample_model_with_rays
    args = cases.los_sampler_case(256, 256, 128, seed=0)
    gpu = sample_model_with_rays("cuda", *args, r_sun_cm=6.957e10)
    cpu = oracle.sample_model_with_rays_cpu(*args, r_sun_cm=6.957e10)
    for k in ("ne", "te", "b", "valid_mask"):
        assert np.array_equal(cpu[k], gpu[k], equal_nan=True), k
    np.testing.assert_allclose(cpu["ds"], gpu["ds"], rtol=1e-6, atol=0)


def test_sampler_edge_cases(oracle, session):
    xg, yg, zg, ne, te, b, r_record, s_arr, ray_start = cases.sampler_fixture(5)
    session.set_field_cubes(xg, yg, zg, ne, te, b)
    # empty inputs
    out = session.sample(np.zeros((0, 4, 3)), np.zeros((0, 4)), np.zeros((4, 3)), 1.0)
    assert out["ne"].shape == (0, 4)
    # all-invalid rays, NaN / inf positions, points exactly on the faces, ragged validity
    r = r_record.copy(); s = s_arr.copy()
    s[:, 0] = 0.0
    s[:, 1] = -1.0
    r[3, 2, 1] = np.nan
    r[4, 3, 2] = np.inf
    r[5, 4] = [1.0, -1.0, 1.0]
    r[6, 5] = [1.0000001, 0.0, 0.0]
    s[::2, 6] = np.nan
    s[:-1, 7] = 0.0
    gpu = session.sample(r, s, ray_start, 6.957e10, fill_ne=-1.0, fill_te=123.0, fill_b=7.0)
    cpu = oracle.sample_model_with_rays_cpu(xg, yg, zg, ne, te, b, r, s, ray_start, 6.957e10, -1.0, 123.0, 7.0)
    for k in ("ne", "te", "b", "valid_mask"):
        assert np.array_equal(cpu[k], gpu[k], equal_nan=True), k
    np.testing.assert_allclose(cpu["ds"], gpu["ds"], rtol=1e-6, atol=0)
    assert not gpu["valid_mask"][:, :2].any() and np.all(gpu["ds"][:, :2] == 0)
    # float64 r_record straight from the integrator (cast to float32 like _as_float32_c)
    gpu64 = session.sample(r.astype(np.float64), s.astype(np.float64), ray_start.astype(np.float64), 6.957e10,
                           -1.0, 123.0, 7.0)
    for k in ("ne", "te", "b", "ds", "valid_mask"):
        assert np.array_equal(gpu64[k], gpu[k], equal_nan=True), k
    with pytest.raises(ValueError):
        session.sample(r[:, :, :2], s, ray_start, 1.0)


def test_sample_traced_equals_host_round_trip(oracle, golden, session):
    kw = cases.trace_case("corona_cs")
    c = synthetic.corona_cube(48, 3.0)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_field_cubes(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"])
    r, s, _ = session.trace(kw["freq_hz"], kw["x_start"], kw["y_start"], kw["z_start"], kw["kvec_in_norm"], kw["dt"],
                            kw["n_steps"], kw["record_stride"], True, 2.0)
    ray_start = np.column_stack([kw["x_start"], kw["y_start"], kw["z_start"]])
    dev = session.sample_traced(ray_start, 6.957e10)
    host = session.sample(r, s, ray_start, 6.957e10)
    for k in ("ne", "te", "b", "ds", "valid_mask", "s"):
        assert np.array_equal(dev[k], host[k], equal_nan=True), k
    # the oracle's sampler on the GPU's own paths: bit for bit
    orc = oracle.sample_model_with_rays_cpu(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"], r, s,
                                            ray_start, 6.957e10)
    for k in ("ne", "te", "b", "valid_mask"):
        assert np.array_equal(dev[k], orc[k], equal_nan=True), k
    np.testing.assert_allclose(dev["ds"], orc["ds"], rtol=1e-6, atol=0)
    # and against the reference's sampler on the reference's own paths (golden): the two path sets differ by
    # |dr| <= 1e-5 R_sun, so validity can only flip on the record at which a ray leaves the cube and the sampled
    # density moves by |grad ln n_e| |dr| (< 1e-4 on this cube)
    g = golden("sampler_on_traced_paths")
    flips = dev["valid_mask"] != g["valid_mask"]
    assert flips.mean() < 1e-3 and flips.sum(axis=0).max() <= 2
    both = dev["valid_mask"] & g["valid_mask"]
    np.testing.assert_allclose(dev["ne"][both], g["ne"][both], rtol=2e-4)


# ------------------------------------------------------------------------------------ GRFF ----
@pytest.mark.parametrize("check", grff_checks.ALL_CHECKS, ids=lambda f: f.__name__)
def test_pyget_mw_analytic(check):
    """The drop-in PyGET_MW symbol bound exactly as the reference binds GRFF's
    (script/resample_with_ray_tracing.py:79-86)."""
    from raytracinggrff_b200 import initGET_MW
    check(initGET_MW())


@pytest.mark.parametrize("cfg", [dict(theta90=True, flag=5, with_b=True), dict(theta90=False, flag=4, with_b=True),
                                 dict(theta90=True, flag=5, with_b=False), dict(theta90=False, flag=0, with_b=True)],
                         ids=["ff_theta90", "grff_theta_var", "ff_noB", "all_on"])
def test_get_mw_slice_matches_oracle(oracle, session, cfg):
    rng = np.random.default_rng(11)
    npix, nz, nf = 257, 77, 4
    P = grff_checks.random_los_batch(rng, npix, nz, **cfg)
    L = np.array([npix, nz, nf, 1, 0, 0], dtype=np.int32)
    R = np.zeros((3, npix), order="F")
    R[0], R[1], R[2] = 2.5e17, 2e8, 0.25
    RL_ref = np.zeros((7, nf, npix), order="F")
    oracle.get_mw_slice(L, R, P, None, None, None, RL_ref)
    RL = np.zeros((7, nf, npix), order="F")
    status = session.get_mw_slice(L, R, P, RL)
    assert np.all(status == 0)
    np.testing.assert_allclose(RL[0], RL_ref[0], rtol=1e-14)
    scale = np.abs(RL_ref[1:]).max(axis=0, keepdims=True) + 1e-300
    assert (np.abs(RL[1:] - RL_ref[1:]) / scale).max() <= TB_RTOL
    np.testing.assert_allclose(RL[1:], RL_ref[1:], rtol=TB_RTOL, atol=1e-12 * float(scale.max()))


def test_get_mw_slice_config2_shape(oracle, session):
    """BASELINE config 2 shape at reduced pixel count: straight LOS, 400 irregular z samples,
    free-free, 4 frequencies from 450 MHz, packed exactly as synthetic_FF_map does (:168-224)."""
    from raytracinggrff_b200 import get_mw_slice
    los = synthetic.straight_los_case(N_pix=48, N_z=400)
    npix, nz, nf = 48 * 48, 400, 4
    ne, te, b, ds = (los[k].reshape(npix, nz) for k in ("Ne_LOS", "Te_LOS", "B_LOS", "ds_LOS"))
    valid = ~(np.isnan(ne) | np.isnan(te) | np.isnan(b))
    P = np.zeros((15, nz, npix), order="F")
    P[4], P[6], P[7] = 90.0, 5, 30
    order = np.argsort(~valid, axis=1, kind="stable")
    keep = np.arange(nz)[None, :] < valid.sum(axis=1)[:, None]
    for m, a in ((0, ds), (1, te), (2, ne), (3, b)):
        P[m] = np.where(keep, np.take_along_axis(a, order, axis=1), 0.0).T
    area = (los["x_coords"][1] - los["x_coords"][0]) ** 2 * 1e4
    L = np.array([npix, nz, nf, 1, 0, 0], dtype=np.int32)
    R = np.zeros((3, npix), order="F")
    R[0], R[1], R[2] = area, 450e6, 0.1
    RL_ref = np.zeros((7, nf, npix), order="F")
    oracle.get_mw_slice(L, R, P, None, None, None, RL_ref)
    RL = np.zeros((7, nf, npix), order="F")
    status = get_mw_slice(L, R, P, 0, 0, 0, RL, tile_pixels=256, heap_bytes=2 << 30)
    assert np.all(status == 0)
    np.testing.assert_allclose(RL[5:], RL_ref[5:], rtol=TB_RTOL, atol=1e-9 * RL_ref[5:].max())
    tb = (RL[5] + RL[6]) * 1e-19 * 2.998e10 ** 2 / (2 * 1.38065e-16 * (RL[0] * 1e9) ** 2) / area * 1.49599e13 ** 2
    assert 5e4 < np.median(tb[0][tb[0] > 0]) < 2.5e6     # quiet-Sun brightness temperatures at 450 MHz


# ---------------------------------------------------------------------- pipeline / fused map ----
def _oracle_chain(oracle, c, N_pix, X_fov, z_obs, freq, dt, n_steps, stride):
    xs, ys, zs, kv = synthetic.ray_launch_geometry(N_pix, X_fov, z_obs)
    r, cs = oracle.ray_trace(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"], freq, xs, ys, zs, kv, dt, n_steps,
                             stride, True, perturb_ratio=2)
    smp = oracle.sample_model_with_rays_cpu(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"], r,
                                            np.array(cs), np.column_stack([xs, ys, zs]), 6.957e10)
    return oracle.emission_from_samples(smp, N_pix, X_fov, freq)


def _cmp_maps(tb, vi, tb_ref, vi_ref):
    assert tb.shape == tb_ref.shape
    assert np.array_equal(tb_ref == 0, tb == 0)
    nz = tb_ref != 0
    rel = np.abs(tb[nz] - tb_ref[nz]) / np.abs(tb_ref[nz])
    assert rel.max() <= TB_RTOL, f"max rel dT_b = {rel.max():.3e}"
    assert np.abs(vi - vi_ref).max() <= TB_RTOL, f"max |d(V/I)| = {np.abs(vi - vi_ref).max():.3e}"


@pytest.mark.parametrize("backend", ["get_mw", "fastgrff", "device", "fused"])
def test_run_ray_tracing_emission_matches_oracle_chain(oracle, backend):
    """The workflow of script/resample_with_ray_tracing.py:295-530 (config-3 physics, 16^2 pixels,
    64^3 cube) through each GRFF backend against oracle trace -> oracle sampler -> oracle GET_MW."""
    from raytracinggrff_b200.workflow import run_ray_tracing_emission
    c = synthetic.corona_cube(64, 3.0)
    args = dict(N_pix=16, X_fov=1.44, freq_hz=75e6, z_observer=3.0, dt=6e-3, n_steps=5000, record_stride=10)
    tb_ref, vi_ref, f_ref = _oracle_chain(oracle, c, 16, 1.44, 3.0, 75e6, 6e-3, 5000, 10)
    res = run_ray_tracing_emission(c, **args, grff_backend=backend, device="cuda", raytrace_device="cuda",
                                   verbose=False)
    assert set(res) >= {"emission_cube", "emission_polVI_cube", "frequencies_Hz", "x_coords", "y_coords"}
    np.testing.assert_array_equal(res["frequencies_Hz"], f_ref)
    assert res["emission_cube"].shape == (16, 16, 1)
    assert (tb_ref > 1e4).mean() > 0.5
    _cmp_maps(res["emission_cube"], res["emission_polVI_cube"], tb_ref, vi_ref)


@pytest.mark.parametrize("backend", ["get_mw", "fastgrff", "device", "fused"])
@pytest.mark.parametrize("s_mode", ["per_step", "cumulative"])
def test_s_input_on_matches_oracle_chain(oracle, backend, s_mode):
    """--s-input-on (script/resample_with_ray_tracing.py:501: Parms[14] = S * area) through every
    backend, with the S of the reference's CPU path (per step) and of its CUDA path (cumulative
    product, gpu_raytrace.py:398-408), against oracle trace -> sampler -> GET_MW fed the same
    Parms[14].  The cumulative S of the oracle is the running product of its per-step ratios."""
    from raytracinggrff_b200.workflow import run_ray_tracing_emission
    c = synthetic.corona_cube(48, 3.0)
    N_pix, X_fov, z_obs, freq, dt, n_steps, stride = 8, 1.44, 3.0, 75e6, 6e-3, 3000, 6
    xs, ys, zs, kv = synthetic.ray_launch_geometry(N_pix, X_fov, z_obs)
    r1, cs1 = oracle.ray_trace(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"], freq, xs, ys, zs, kv, dt, n_steps,
                               1, True, perturb_ratio=2)
    s_all = np.array(cs1)
    if s_mode == "cumulative":
        s_all = np.cumprod(s_all, axis=0)
    smp = oracle.sample_model_with_rays_cpu(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"],
                                            r1[::stride], s_all[::stride], np.column_stack([xs, ys, zs]), 6.957e10)
    tb_ref, vi_ref, _ = oracle.emission_from_samples(smp, N_pix, X_fov, freq, s_input_on=True)
    tb_off, _, _ = oracle.emission_from_samples(smp, N_pix, X_fov, freq, s_input_on=False)
    res = run_ray_tracing_emission(c, N_pix=N_pix, X_fov=X_fov, freq_hz=freq, z_observer=z_obs, dt=dt, n_steps=n_steps,
                                   record_stride=stride, grff_backend=backend, s_input_on=True, s_mode=s_mode,
                                   verbose=False)
    _cmp_maps(res["emission_cube"], res["emission_polVI_cube"], tb_ref, vi_ref)
    # the S input changes the map (by the pencil's magnification in cumulative mode)
    on = tb_ref != 0
    assert np.abs(tb_ref[on] / tb_off[on] - 1).max() > (1e-3 if s_mode == "cumulative" else 1e-6)


def test_render_map_multi_frequency_and_orders(oracle, session):
    """Fused map at three frequencies with per-frequency presets; record order (reference behaviour)
    and reversed order (far end first) against the oracle fed the same / reversed voxel lists;
    theta from the B vector against a numpy restatement of the angle."""
    c = synthetic.corona_cube(64, 3.0, active_region=True)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_field_cubes(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"], c["bx"], c["by"], c["bz"])
    N_pix, X_fov = 12, 1.2
    xs, ys, zs, kv = synthetic.ray_launch_geometry(N_pix, X_fov, 3.0)
    area = (2 * X_fov / N_pix * 6.957e10) ** 2
    freqs = [75e6, 150e6, 400e6]
    fps = [dict(freq_hz=f, **synthetic.frequency_scaled_params(f)) for f in freqs]
    tb, vi, stats = session.render_map(xs, ys, zs, fps, kvec_in_norm=kv, pixel_area_cm2=area)
    assert tb.shape == (3, N_pix * N_pix)
    assert stats["nominal_ray_steps"] == sum(p["n_steps"] for p in fps) * xs.size
    assert 0 < stats["active_ray_steps"] <= stats["nominal_ray_steps"]
    # the counters of the fused kernel (derived from the step at which each ray froze) equal the trace kernel's
    act_sum = pen_sum = 0
    for p in fps:
        _, _, act = session.trace(p["freq_hz"], xs, ys, zs, kv, p["dt"], p["n_steps"], p["record_stride"], True, 2.0,
                                  fetch=False)
        act_sum += act
    assert stats["active_ray_steps"] == act_sum
    assert 0 < stats["pencil_steps"] <= stats["active_ray_steps"] and stats["valid_samples"] <= stats["pencil_steps"]
    for i, p in enumerate(fps):
        tb_ref, vi_ref, _ = _oracle_chain(oracle, c, N_pix, X_fov, 3.0, p["freq_hz"], p["dt"], p["n_steps"],
                                          p["record_stride"])
        _cmp_maps(tb[i].reshape(N_pix, N_pix, 1), vi[i].reshape(N_pix, N_pix, 1), tb_ref, vi_ref)
    # reversed voxel order == oracle GET_MW on the reversed compacted list
    p = fps[1]
    tb_r, vi_r, _ = session.render_map(xs, ys, zs, [p], kvec_in_norm=kv, pixel_area_cm2=area, voxel_order=1)
    r, cs = oracle.ray_trace(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"], p["freq_hz"], xs, ys, zs, kv,
                             p["dt"], p["n_steps"], p["record_stride"], True, perturb_ratio=2)
    smp = oracle.sample_model_with_rays_cpu(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"], r,
                                            np.array(cs), np.column_stack([xs, ys, zs]), 6.957e10)
    rev = {k: v[::-1].copy() for k, v in smp.items()}
    tb_ref, vi_ref, _ = oracle.emission_from_samples(rev, N_pix, X_fov, p["freq_hz"])
    _cmp_maps(tb_r[0].reshape(N_pix, N_pix, 1), vi_r[0].reshape(N_pix, N_pix, 1), tb_ref, vi_ref)


def test_render_map_theta_from_bvec_gr_on(oracle, session):
    """GR+FF with theta from B.t along the ray (beyond the reference, which fixes theta = 90 deg):
    the oracle is fed Parms built in numpy from the oracle's own paths — B vector sampled with the
    sampler's float32 arithmetic, theta = acos(-B.d/|B||d|) with d the step between consecutive valid
    float32 samples, |B| = |B vector|, flag 4 (GR and FF on), s_max 30."""
    c = synthetic.corona_cube(64, 3.0, active_region=True)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_field_cubes(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"], c["bx"], c["by"], c["bz"])
    N_pix, X_fov, freq = 12, 1.2, 1.0e9
    p = synthetic.frequency_scaled_params(freq)
    p["record_stride"] = 4
    xs, ys, zs, kv = synthetic.ray_launch_geometry(N_pix, X_fov, 3.0)
    area = (2 * X_fov / N_pix * 6.957e10) ** 2
    tb, vi, _ = session.render_map(xs, ys, zs, [dict(freq_hz=freq, **p)], kvec_in_norm=kv, pixel_area_cm2=area,
                                   em_flag=4, s_max=30, use_bvec=True)
    r, cs = oracle.ray_trace(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"], freq, xs, ys, zs, kv, p["dt"],
                             p["n_steps"], p["record_stride"], True, perturb_ratio=2)
    ray_start = np.column_stack([xs, ys, zs])
    g3 = (c["x_grid"], c["y_grid"], c["z_grid"])
    smp = oracle.sample_model_with_rays_cpu(*g3, c["ne"], c["te"], c["b"], r, np.array(cs), ray_start, 6.957e10)
    bv = oracle.sample_model_with_rays_cpu(*g3, c["bx"], c["by"], c["bz"], r, np.array(cs), ray_start, 6.957e10,
                                           fill_ne=0.0, fill_te=0.0, fill_b=0.0)
    pos = r.astype(np.float32).astype(np.float64)
    start32 = ray_start.astype(np.float32).astype(np.float64)
    tb_ref = np.zeros(N_pix * N_pix)
    vi_ref = np.zeros(N_pix * N_pix)
    n_gr = 0
    for q in range(N_pix * N_pix):
        valid = smp["valid_mask"][:, q]
        idx = np.flatnonzero(valid)
        if idx.size == 0:
            continue
        prev = np.vstack([start32[q][None], pos[idx[:-1], q]])
        d = pos[idx, q] - prev
        B = np.stack([bv["ne"][idx, q], bv["te"][idx, q], bv["b"][idx, q]], axis=1).astype(np.float64)
        b2 = (B * B).sum(1)
        dn2 = (d * d).sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            cth = np.clip(-(B * d).sum(1) / np.sqrt(b2 * dn2), -1.0, 1.0)
        cth = np.where((b2 > 0) & (dn2 > 0), cth, 6.123233995736766e-17)
        P = np.zeros((15, idx.size), order="F")
        P[0] = smp["ds"][idx, q]; P[1] = smp["te"][idx, q]; P[2] = smp["ne"][idx, q]; P[3] = np.sqrt(b2)
        P[4] = np.degrees(np.arccos(cth)); P[6] = 4; P[7] = 30
        RL = np.zeros((7, 1), order="F")
        assert oracle.get_mw(np.array([idx.size, 1, 0, 0, 0], dtype=np.int32), np.array([area, freq, 0.0]), P,
                             None, None, None, RL) == 0
        conv = (1e-19 * 2.998e10 ** 2 / (2.0 * 1.38065e-16 * (RL[0, 0] * 1e9) ** 2) / area) * 1.49599e13 ** 2
        tb_ref[q] = (RL[5, 0] + RL[6, 0]) * conv
        vi_ref[q] = (RL[5, 0] - RL[6, 0]) / (RL[5, 0] + RL[6, 0] + 1e-30)
        bres = 3.57238675287821e-07 * freq / np.arange(2, 31)
        n_gr += int(np.any((P[3].min() < bres) & (bres < P[3].max())))
    assert n_gr > 0                      # the active region puts gyroresonance layers on some rays
    assert np.abs(vi_ref).max() > 1e-3   # and polarises the map
    _cmp_maps(tb[0].reshape(N_pix, N_pix, 1), vi[0].reshape(N_pix, N_pix, 1), tb_ref.reshape(N_pix, N_pix, 1),
              vi_ref.reshape(N_pix, N_pix, 1))


def test_workflow_command_line(tmp_path, oracle):
    """The command line of script/resample_with_ray_tracing.py:652-730 (same flags) on a spherical model
    file: cubes built on the GPU, map written with the reference's npz keys (:533-540), --consider-beam
    adds the convolved map; --device cpu is refused (no CPU path)."""
    from raytracinggrff_b200 import cubes, workflow
    model = synthetic.spherical_corona(40, 30, 48, r_max=6.0)
    path = tmp_path / "model.npz"
    cubes.save_spherical_model(path, model)
    out = tmp_path / "map.npz"
    argv = ["-m", str(path), "-n", "12", "--grid-n", "48", "--n-steps", "2500", "--record-stride", "10", "-o", str(out),
            "--device", "cuda", "--raytrace-device", "cuda", "--grff-backend", "fused", "--consider-beam",
            "--beam-fwhm", "0.3", "--phi0-offset", "-30", "-q"]
    res = workflow.main(argv)
    z = np.load(out)
    assert set(z.files) >= {"emission_cube", "emission_polVI_cube", "frequencies_Hz", "x_coords", "y_coords",
                            "emission_map_beam"}
    assert z["emission_cube"].shape == (12, 12, 1) and z["frequencies_Hz"][0] == 75e6
    assert (z["emission_cube"] > 1e4).mean() > 0.4
    # the same map through the per-pixel GET_MW backend
    res2 = workflow.main(argv[:-8] + ["--grff-backend", "get_mw", "-q"])
    _cmp_maps(res["emission_cube"], res["emission_polVI_cube"], res2["emission_cube"], res2["emission_polVI_cube"])
    beam = res["emission_cube"][:, :, 0].copy()
    beam[beam == 0] = np.nan
    np.testing.assert_allclose(z["emission_map_beam"], oracle.gaussian_filter(beam, 0.3 / 2.88 * 12), rtol=1e-12,
                               equal_nan=True)
    with pytest.raises(RuntimeError):
        workflow.main(argv[:-1] + ["--device", "cpu", "-q"])
    with pytest.raises(FileNotFoundError):
        workflow.main(["-m", str(tmp_path / "missing.npz"), "-q"])


def test_trace_steps_longer_than_a_cell(oracle, session):
    """Low-frequency presets on a fine cube make a step (plus the pencil offset) span more than one
    cell (BASELINE config 5 at 20-24 MHz: 1.1 cells).  The FP32 stepper handles whole-cell shifts of
    any size; the face margin grows with the step (StepConst.margin)."""
    n = 160
    g = np.linspace(-1.2, 1.2, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
    w = 2 * np.pi * 45e6 * np.exp(-(r - 0.3)) * (1 + 0.2 * np.sin(3 * X) * np.cos(2 * Y))
    session.set_omega_cube(w, g, g, g)
    rng = np.random.default_rng(21)
    m = 64
    xs, ys = rng.uniform(-0.9, 0.9, m), rng.uniform(-0.9, 0.9, m)
    zs = np.full(m, 1.2)
    zs[:4] = [1.2000001, 1.19, 1.0, 0.5]
    kv = np.column_stack([rng.normal(scale=0.15, size=m), rng.normal(scale=0.15, size=m), -np.ones(m)])
    kv /= np.linalg.norm(kv, axis=1, keepdims=True)
    for dt, steps in ((0.0134, 500), (0.03, 260)):       # 1.15 and 2.6 cells per (step + pencil)
        cells = 3 * dt * (2.998e10 / 6.96e10) * (n - 1) / 2.4
        assert cells > 1.1
        rr, ss, act = session.trace(80e6, xs, ys, zs, kv, dt, steps, 7, True, 2.0)
        r_ref, cs_ref, act_ref = oracle.ray_trace(w, g, g, g, 80e6, xs, ys, zs, kv, dt, steps, 7, True, perturb_ratio=2,
                                                  return_active=True)
        _cmp_paths(rr, ss, r_ref, np.array(cs_ref), lo=np.full(3, -1.2), hi=np.full(3, 1.2), step_len=1.01 * dt * 0.43075)
        assert abs(act - act_ref) <= 2


def test_render_map_properties_at_full_config4_size(session):
    """BASELINE config 4 geometry at full size (512^2 pixels, 256^3 cube) through the fused kernel, checked
    through size-independent properties: bit-identical repeats (no race, no order dependence), the
    mirror symmetry x -> -x of the axisymmetric corona, brightness temperatures bounded by the hottest
    plasma on the path, weak circular polarisation at theta = 90, and
    the image tiling (ray_order) not changing a single bit."""
    c = synthetic.corona_cube(256, 3.0)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_field_cubes(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"])
    n = 512
    xs, ys, zs, kv = synthetic.ray_launch_geometry(n, 1.44, 3.0)
    area = (2 * 1.44 / n * 6.957e10) ** 2
    fps = [dict(freq_hz=f, **synthetic.frequency_scaled_params(f)) for f in (75e6, 400e6)]
    tb, vi, st = session.render_map(xs, ys, zs, fps, pixel_area_cm2=area, image_shape=(n, n))
    tb2, vi2, st2 = session.render_map(xs, ys, zs, fps, pixel_area_cm2=area, image_shape=(n, n))
    assert np.array_equal(tb, tb2) and np.array_equal(vi, vi2) and st == st2
    tb3, vi3, _ = session.render_map(xs, ys, zs, fps, pixel_area_cm2=area)            # row-major threads
    assert np.array_equal(tb, tb3) and np.array_equal(vi, vi3)
    img = tb.reshape(2, n, n)
    assert np.isfinite(img).all() and img.min() >= 0.0
    assert img.max() <= 1.0001 * float(c["te"].max())
    assert (img[0] > 1e5).mean() > 0.3                     # the disc is bright at 75 MHz
    # mirror symmetry at 75 MHz, where every ray turns around in the smooth corona (at 400 MHz some rays
    # reach the r = 1 density discontinuity and are chaotic: test_rays_through_the_density_discontinuity)
    mirror = img[0][:, ::-1]
    on = (img[0] > 1e3) & (mirror > 1e3)
    rel = np.abs(img[0] - mirror)[on] / img[0][on]
    assert rel.max() < 1e-5, rel.max()
    # theta = 90 deg with |B| > 0: the X mode (taken as R at cos(theta) >= 0) is the more opaque one
    assert 0 < np.abs(vi).max() < 0.05
    assert st["nominal_ray_steps"] == n * n * sum(p["n_steps"] for p in fps)


def test_rays_through_the_density_discontinuity(oracle, session):
    """Where the tolerance on the paths cannot hold for ANY implementation that is not bit-identical: at
    400 MHz on the config-4 cube some rays reach the solar surface, where the model's density drops to
    zero across one cell (fill value inside r < 1, script/resample_with_ray_tracing.py:269-279).  A ray
    that crosses that jump at grazing incidence amplifies a 1e-9 perturbation (here: the float32 storage
    of the cube) to O(1) R_sun.  They are few, they are exactly the rays that dive below r = 1, and every
    other ray is within the tolerance (median deviation < 5e-7 R_sun)."""
    c = synthetic.corona_cube(256, 3.0)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    n = 512
    xs, ys, zs, kv = synthetic.ray_launch_geometry(n, 1.44, 3.0)
    sel = (np.arange(0, n, 16)[:, None] * n + np.arange(0, n, 16)[None, :]).ravel()
    xs, ys, zs, kv = xs[sel], ys[sel], zs[sel], kv[sel]
    f = 400e6
    p = synthetic.frequency_scaled_params(f)
    r, s, _ = session.trace(f, xs, ys, zs, kv, p["dt"], p["n_steps"], 8, True, 2.0)
    r_ref, _ = oracle.ray_trace(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"], f, xs, ys, zs, kv, p["dt"],
                                p["n_steps"], 8, True, perturb_ratio=2)
    inside = np.all(np.abs(r_ref) <= 3.0, axis=2)
    dev = np.where(inside, np.abs(r - r_ref).max(axis=2), 0.0).max(axis=0)
    closest = np.where(inside, np.linalg.norm(r_ref, axis=2), 9.0).min(axis=0)
    bad = dev > POS_TOL
    assert bad.mean() < 0.03, f"{bad.sum()} of {bad.size} rays deviate"
    assert np.all(closest[bad] < 1.0), "only rays that dive below the surface may deviate"
    assert np.median(dev) < 5e-7 and np.quantile(dev, 0.95) < 2e-6


def test_trace_backwards_in_time(oracle, session):
    """dt < 0 (the reference takes any float dt): parity with the oracle, and a forward trace followed
    by a backward trace from its end point (k reversed in time = same k, negative dt) retraces the path."""
    c = synthetic.corona_cube(48, 3.0)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    rng = np.random.default_rng(4)
    m = 32
    xs, ys, zs = rng.uniform(-1.0, 1.0, m), rng.uniform(-1.0, 1.0, m), rng.uniform(1.6, 2.4, m)
    kv = np.column_stack([rng.normal(scale=0.3, size=m), rng.normal(scale=0.3, size=m), np.ones(m)])
    kv /= np.linalg.norm(kv, axis=1, keepdims=True)
    r, s, _ = session.trace(90e6, xs, ys, zs, kv, -5e-3, 300, 5, True, 2.0)
    r_ref, cs_ref = oracle.ray_trace(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"], 90e6, xs, ys, zs, kv, -5e-3,
                                     300, 5, True, perturb_ratio=2)
    _cmp_paths(r, s, r_ref, np.array(cs_ref), lo=np.full(3, -3.0), hi=np.full(3, 3.0), step_len=1.01 * 5e-3 * 0.43075)
    assert np.nanmax(np.abs(r[-1] - np.column_stack([xs, ys, zs]))) > 0.1      # the rays did move


def test_render_map_many_frequencies_order_and_chunks(session):
    """More frequencies than fit one launch (16 per launch, dispatched longest first): every row of the
    result must equal the single-frequency render of that frequency, bit for bit."""
    c = synthetic.corona_cube(48, 3.0, active_region=True)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_field_cubes(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"], c["bx"], c["by"], c["bz"])
    xs, ys, zs, kv = synthetic.ray_launch_geometry(6, 1.2, 3.0)
    area = (2 * 1.2 / 6 * 6.957e10) ** 2
    rng = np.random.default_rng(8)
    freqs = rng.permutation(np.geomspace(60e6, 900e6, 19))          # unsorted on purpose
    fps = [dict(freq_hz=float(f), dt=6e-3 * (100e6 / f) ** 0.5, n_steps=int(600 + 40 * i), record_stride=1 + i % 4)
           for i, f in enumerate(freqs)]
    kw = dict(kvec_in_norm=kv, pixel_area_cm2=area, em_flag=4, use_bvec=True)
    tb, vi, st = session.render_map(xs, ys, zs, fps, **kw)
    assert tb.shape == (19, 36) and st["nominal_ray_steps"] == 36 * sum(p["n_steps"] for p in fps)
    act = 0
    for i, p in enumerate(fps):
        tb1, vi1, st1 = session.render_map(xs, ys, zs, [p], **kw)
        assert np.array_equal(tb[i], tb1[0]) and np.array_equal(vi[i], vi1[0]), i
        act += st1["active_ray_steps"]
    assert act == st["active_ray_steps"]
    assert (tb > 0).any()
