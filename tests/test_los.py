"""Straight-LOS workflow (BASELINE config 2 shape): resample_MAS + SyntheticFF drop-ins."""
import numpy as np
import pytest

from raytracinggrff_b200 import los, synthetic


def test_z_grid_matches_reference_formula():
    zc, dz = los.z_grid(400, 3e-4)
    i = np.arange(400)
    np.testing.assert_array_equal(dz, 3e-4 * (1 + (5 * i / 400) ** 2.5))
    np.testing.assert_array_equal(zc, np.cumsum(dz))
    z, d = los.z_grid(5, 0.0, variable_spacing_z=False)
    np.testing.assert_allclose(z, [0, 1, 2, 3, 4]); np.testing.assert_allclose(d, [0, 1, 1, 1, 1])
    with pytest.raises(ValueError, match="extremely large"):
        los.resample_MAS({}, 4, (-1, 1), (-1, 1), 10, 7e4)


@pytest.mark.gpu
def test_resample_mas_and_synthetic_ff_match_oracle(oracle, session, tmp_path):
    from oracle import oracle_cubes as oc
    model = synthetic.spherical_corona(64, 48, 64, r_max=8.0, active_region=True)
    kw = dict(N_pix=14, X_range=(-1.44, 1.44), Y_range=(-1.44, 1.44), N_z=120, dz0=1e-3)
    ref = oc.resample_MAS(model, phi0_offset=24.0, **kw)
    out = tmp_path / "LOS_data.npz"
    got = los.resample_MAS(model, out_path=out, phi0_offset=24.0, context=session.ctx, **kw)
    saved = np.load(out)
    assert set(saved.files) == {"Ne_LOS", "Te_LOS", "B_LOS", "ds_LOS", "x_coords", "y_coords", "z_coords"}
    for k in ("Ne_LOS", "Te_LOS", "B_LOS", "ds_LOS", "x_coords", "y_coords", "z_coords"):
        assert got[k].shape == ref[k].shape, k
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref[k])), k
        np.testing.assert_allclose(np.nan_to_num(got[k]), np.nan_to_num(ref[k]), rtol=1e-9, atol=1e-9 * np.nanmax(np.abs(ref[k])))
    # SyntheticFF against the reference's per-pixel loop restated in the oracle
    area = ((ref["x_coords"][1] - ref["x_coords"][0]) / (6.957e10 * 1e-2) * 6.957e10) ** 2
    tb_ref, vi_ref, f_ref = oracle.emission_from_los(ref["Ne_LOS"], ref["Te_LOS"], ref["B_LOS"], ref["ds_LOS"], area, 450e6, 4, 0.1)
    res = los.SyntheticFF(out, 450e6, 4, 0.1, fname_output=tmp_path / "ff", session=session)
    assert set(np.load(str(tmp_path / "ff") + ".npz").files) == {"emission_cube", "emission_polVI_cube", "frequencies_Hz", "x_coords", "y_coords"}
    np.testing.assert_array_equal(res["frequencies_Hz"], f_ref)
    assert res["emission_cube"].shape == (14, 14, 4)
    nz = tb_ref > 0
    assert nz.mean() > 0.9
    np.testing.assert_allclose(res["emission_cube"][nz], tb_ref[nz], rtol=1e-4)
    np.testing.assert_allclose(res["emission_polVI_cube"][nz], vi_ref[nz], atol=1e-4)
    assert 1e5 < np.median(res["emission_cube"][..., 0]) < 2e6
