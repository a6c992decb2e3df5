"""Cube builder: spherical model -> xyz cubes.  CPU part: host helpers and
the oracle restatement; GPU part: the CUDA resampler / composer against the oracle."""
import numpy as np
import pytest

from raytracinggrff_b200 import synthetic
from raytracinggrff_b200 import cubes


def test_cart_to_sph_matches_oracle_and_reference_convention():
    from oracle import oracle_cubes as oc
    rng = np.random.default_rng(0)
    x, y, z = rng.normal(size=(3, 200))
    for off in (0.0, 24.0, -100.0):
        a = cubes.cart_to_sph(x, y, z, off)
        b = oc.cart_to_sph(x, y, z, off)
        for u, v in zip(a, b):
            np.testing.assert_array_equal(u, v)
    r, colat, lon = cubes.cart_to_sph(np.array([0.0]), np.array([-1.0]), np.array([0.0]))
    assert r[0] == 1 and colat[0] == pytest.approx(np.pi / 2) and lon[0] == pytest.approx(1.5 * np.pi)


def test_spherical_model_reproduces_the_analytic_corona():
    from oracle import oracle_cubes as oc
    m = synthetic.spherical_corona(90, 70, 96, active_region=False)
    assert m["br"].r.size == m["rho"].r.size + 1 and m["bt"].lat.size == m["rho"].lat.size + 1   # staggered
    g = np.linspace(-3.0, 3.0, 20)
    c = oc.compose_cubes(m, g, g, g)
    ref = synthetic.corona_cube(20, 3.0)
    out = ref["ne"] > 0
    assert np.median(np.abs(c["ne"][out] - ref["ne"][out]) / ref["ne"][out]) < 5e-3
    assert np.median(np.abs(c["b"][out] - ref["b"][out]) / ref["b"][out]) < 5e-3
    assert np.all(c["ne"][~out] == 0) and np.all(c["te"][~out] == 1e4) and np.all(c["omega_pe"][~out] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("phi0", (0.0, 24.0))
def test_resample_to_xyz_cube_matches_oracle(session, phi0):
    from oracle import oracle_cubes as oc
    m = synthetic.spherical_corona(48, 40, 50, active_region=True)
    xg = np.linspace(-3.0, 3.0, 37); yg = np.linspace(-2.0, 2.5, 29); zg = np.linspace(-3.0, 1.0, 33)
    for name, fill in (("rho", 0.0), ("te", None), ("br", 0.0), ("bp", -7.0)):
        gpu = cubes.resample_to_xyz_cube(m, name, xg, yg, zg, phi0_offset=phi0, fill_nan=fill, context=session.ctx)
        ref = oc.resample_to_xyz_cube(m[name], xg, yg, zg, phi0_offset=phi0, fill_nan=fill)
        assert gpu.shape == ref.shape and gpu.dtype == np.float64
        assert np.array_equal(np.isnan(gpu), np.isnan(ref)), name
        scale = np.nanmax(np.abs(ref))
        assert np.nanmax(np.abs(gpu - ref)) <= 1e-10 * scale, name
    with pytest.raises(ValueError):
        bad = cubes.SphericalVariable(m["rho"].data, m["rho"].phi[::-1], m["rho"].lat, m["rho"].r)
        cubes.resample_to_xyz_cube({"rho": bad}, "rho", xg, yg, zg, context=session.ctx)


@pytest.mark.gpu
def test_cubes_composed_on_device_drive_the_same_rays(oracle, session):
    """set_model_from_spherical == uploading the oracle-composed cubes: same paths, same samples,
    same map (script/resample_with_ray_tracing.py:263-293 without the cubes visiting the host)."""
    from oracle import oracle_cubes as oc
    m = synthetic.spherical_corona(64, 48, 64, active_region=True)
    g = np.linspace(-3.0, 3.0, 48)
    c = oc.compose_cubes(m, g, g, g, phi0_offset=24.0)
    xs, ys, zs, kv = synthetic.ray_launch_geometry(10, 1.3, 3.0)
    start = np.column_stack([xs, ys, zs])
    area = (2 * 1.3 / 10 * 6.957e10) ** 2
    session.set_omega_cube(c["omega_pe"], g, g, g)
    session.set_field_cubes(g, g, g, c["ne"], c["te"], c["b"])
    r_a, s_a, _ = session.trace(75e6, xs, ys, zs, kv, 6e-3, 2500, 10, True, 2.0)
    smp_a = session.sample_traced(start, 6.957e10)
    tb_a, vi_a = session.emission_traced(area, 75e6)
    session.set_model_from_spherical(m, g, g, g, phi0_offset=24.0, want_bvec=True)
    r_b, s_b, _ = session.trace(75e6, xs, ys, zs, kv, 6e-3, 2500, 10, True, 2.0)
    smp_b = session.sample_traced(start, 6.957e10)
    tb_b, vi_b = session.emission_traced(area, 75e6)
    assert np.nanmax(np.abs(r_a - r_b)) < 1e-7
    assert np.array_equal(smp_a["valid_mask"], smp_b["valid_mask"])
    v = smp_a["valid_mask"]
    for k in ("ne", "te", "b"):
        np.testing.assert_allclose(smp_a[k][v], smp_b[k][v], rtol=2e-6)
    nz = tb_a != 0
    assert nz.any()
    np.testing.assert_allclose(tb_b[nz], tb_a[nz], rtol=1e-5)
    # the Cartesian B vector built on the device has the magnitude of the |B| cube
    tb_v, vi_v, _ = session.render_map(xs, ys, zs, [(75e6, 6e-3, 2500, 10)], kvec_in_norm=kv, pixel_area_cm2=area,
                                       em_flag=5, use_bvec=True)
    assert np.all(np.isfinite(tb_v)) and (tb_v[0] > 1e4).mean() > 0.3
