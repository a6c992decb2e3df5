"""CPU: the oracle against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  The integrator restatement is bit-exact with
build_rays.ray_trace on these cases, the sampler restatement with _sample_model_with_rays_cpu."""
import numpy as np
import pytest

import cases
from raytracinggrff_b200 import synthetic


@pytest.mark.parametrize("name", cases.TRACE_CASES)
def test_oracle_ray_trace_matches_reference(oracle, golden, name):
    kw = cases.trace_case(name)
    g = golden(f"trace_{name}")
    r, cs = oracle.ray_trace(**kw)
    ref = g["r_record"]
    assert r.shape == ref.shape and r.dtype == np.float64
    # n_rec = ceil(n_steps / stride)  (build_rays.py:241-244)
    assert r.shape[0] == -(-kw["n_steps"] // kw["record_stride"])
    assert np.array_equal(np.isnan(r), np.isnan(ref))
    np.testing.assert_array_equal(np.nan_to_num(r), np.nan_to_num(ref))
    if kw["trace_crosssections"]:
        s, sref = np.array(cs), g["s_record"]
        assert np.array_equal(np.isnan(s), np.isnan(sref))
        np.testing.assert_array_equal(np.nan_to_num(s), np.nan_to_num(sref))
    else:
        assert cs == []


def test_oracle_gradient_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    f = rng.normal(size=(7, 5, 9))
    for axis, h in enumerate((0.3, 0.7, 1.1)):
        np.testing.assert_array_equal(oracle.gradient(f, h, axis), np.gradient(f, h, axis=axis))


@pytest.mark.parametrize("seed", (1, 2, 3))
def test_oracle_sampler_matches_reference_fixture(oracle, golden, seed):
    args = cases.sampler_fixture(seed)
    out = oracle.sample_model_with_rays_cpu(*args, r_sun_cm=1.0)
    g = golden(f"sampler_fixture_seed{seed}")
    for k in ("ne", "te", "b", "ds", "valid_mask", "s"):
        assert out[k].dtype == g[k].dtype, k
        assert np.array_equal(out[k], g[k], equal_nan=True), k


def test_oracle_sampler_c1_and_traced_paths(oracle, golden):
    args = cases.los_sampler_case(24, 48, 40, seed=0)
    out = oracle.sample_model_with_rays_cpu(*args, r_sun_cm=6.957e10)
    g = golden("sampler_c1_small")
    for k in ("ne", "te", "b", "ds", "valid_mask"):
        assert np.array_equal(out[k], g[k], equal_nan=True), k
    kw = cases.trace_case("corona_cs")
    t = golden("trace_corona_cs")
    c = synthetic.corona_cube(48, 3.0)
    ray_start = np.column_stack([kw["x_start"], kw["y_start"], kw["z_start"]])
    out = oracle.sample_model_with_rays_cpu(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"],
                                            t["r_record"], t["s_record"], ray_start, r_sun_cm=6.957e10)
    g = golden("sampler_on_traced_paths")
    for k in ("ne", "te", "b", "ds", "valid_mask", "s"):
        assert np.array_equal(out[k], g[k], equal_nan=True), k


# The reference's own tests (tests/test_gpu_raytrace.py:47-88), run against the oracle.
def test_reference_test_linear_field_accuracy(oracle):
    xg, yg, zg, ne, te, b, r_record, s_arr, ray_start = cases.sampler_fixture(seed=1)
    out = oracle.sample_model_with_rays_cpu(xg, yg, zg, ne, te, b, r_record, s_arr, ray_start, r_sun_cm=1.0)
    valid = out["valid_mask"]
    inb = ((r_record[..., 0] >= xg[0]) & (r_record[..., 0] <= xg[-1]) & (r_record[..., 1] >= yg[0])
           & (r_record[..., 1] <= yg[-1]) & (r_record[..., 2] >= zg[0]) & (r_record[..., 2] <= zg[-1]))
    mask = valid & inb
    expected_ne = r_record[..., 0] + r_record[..., 1] + r_record[..., 2]
    np.testing.assert_allclose(out["ne"][mask], expected_ne[mask], rtol=2e-5, atol=2e-5)
    oob = valid & ~inb
    assert np.any(oob)
    np.testing.assert_allclose(out["ne"][oob], 0.0)
    np.testing.assert_allclose(out["te"][oob], 1e4)
    np.testing.assert_allclose(out["b"][oob], 0.0)


def test_reference_test_valid_mask_and_ds_shape(oracle):
    xg, yg, zg, ne, te, b, r_record, s_arr, ray_start = cases.sampler_fixture(seed=2)
    out = oracle.sample_model_with_rays_cpu(xg, yg, zg, ne, te, b, r_record, s_arr, ray_start, r_sun_cm=1.0)
    for k in ("ne", "te", "b", "ds", "valid_mask"):
        assert out[k].shape == s_arr.shape
    assert np.all(~out["valid_mask"][::9, ::7])
    assert np.all(out["ds"] >= 0.0)


# Invariants the reference does not pin, checked on the oracle.
def test_vacuum_rays_are_straight(oracle):
    g = np.linspace(-2.0, 2.0, 17)
    w = np.zeros((17, 17, 17))
    xs = np.array([0.1, -0.7, 1.3]); ys = np.array([0.2, 0.5, -1.1]); zs = np.full(3, 1.9)
    kv = np.array([[0, 0, -1.0], [0.6, 0, -0.8], [0, -0.6, -0.8]])
    dt, n = 5e-3, 400
    r, cs = oracle.ray_trace(w, g, g, g, 80e6, xs, ys, zs, kv, dt, n, record_stride=1, trace_crosssections=True)
    t = dt * np.arange(1, n + 1)
    expect = np.stack([xs, ys, zs], 1)[None] + oracle.C_R * t[:, None, None] * kv[None]
    inside = np.all(np.abs(expect) <= 2.0, axis=2)
    assert inside.sum() > 600
    np.testing.assert_allclose(r[inside], expect[inside], rtol=0, atol=1e-12)
    s = np.array(cs)
    np.testing.assert_allclose(s[inside][5:], 1.0, rtol=0, atol=1e-9)      # S ~ 1 in vacuum


def test_frozen_after_exit_and_hamiltonian(oracle):
    kw = cases.trace_case("corona_cs")
    r, cs = oracle.ray_trace(**kw)
    ext = kw["x_grid"][-1]
    out = np.any(np.abs(r) > ext, axis=2)
    # once outside, the record repeats forever and S is NaN
    first_out = np.argmax(out, axis=0)
    for ray in np.flatnonzero(out.any(axis=0))[:16]:
        i = first_out[ray]
        assert np.all(r[i:, ray] == r[i, ray])
        assert np.all(np.isnan(np.array(cs)[i + 1:, ray]))
