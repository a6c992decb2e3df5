"""Frequency-sweep driver."""
import json
from pathlib import Path

import numpy as np
import pytest

from raytracinggrff_b200 import sweep, synthetic

GOLDEN = Path(__file__).resolve().parent / "golden" / "select_params.json"


def test_select_params_matches_reference_golden():
    ref = json.load(open(GOLDEN))
    assert len(ref) >= 12
    for f_repr, want in ref.items():
        got = sweep.select_params(float(f_repr))
        assert set(got) == set(want)
        for k, v in want.items():
            assert got[k] == v, (f_repr, k, got[k], v)
            assert type(got[k]) is type(v)


def test_sweep_argument_checks(tmp_path):
    with pytest.raises(ValueError, match="start-from-idx"):
        sweep.tb_spectra({}, tmp_path, n_freq=4, start_from_idx=4, session=object())
    with pytest.raises(RuntimeError, match="CUDA-only"):
        sweep.main(["--device", "cpu", "--out-dir", str(tmp_path)])


@pytest.mark.gpu
def test_tb_spectra_files_manifest_and_resume(session, tmp_path):
    model = synthetic.spherical_corona(64, 48, 64, r_max=8.0)
    kw = dict(N_pix=12, fmin_mhz=40.0, fmax_mhz=400.0, n_freq=4, phi0_offset=-140.0, session=session, max_grid_n=64)
    rows = sweep.tb_spectra(model, tmp_path, **kw)
    assert [r[0] for r in rows] == [0, 1, 2, 3]
    man = (tmp_path / "TbSpectra_manifest.txt").read_text().splitlines()
    assert man[0] == "# idx freq_hz npz_path png_path" and len(man) == 5
    for i, f, path in rows:
        assert Path(path).name == f"raytrace_{i:02d}_{f/1e6:08.3f}MHz.npz"
        d = np.load(path)
        assert set(d.files) == {"emission_cube", "emission_polVI_cube", "frequencies_Hz", "x_coords", "y_coords"}
        assert d["emission_cube"].shape == (12, 12, 1) and d["frequencies_Hz"][0] == f
        p = sweep.select_params(f)
        np.testing.assert_allclose(d["x_coords"], np.linspace(-p["x_fov"], p["x_fov"], 12) * 6.957e8)
        tb = d["emission_cube"][:, :, 0]
        assert np.all(np.isfinite(tb)) and 1e4 < tb.max() < 3e6
    first = np.load(rows[0][2])["emission_cube"].copy()
    # resume: indices below start_from_idx are kept, the others recomputed identically
    np.savez_compressed(rows[0][2], **{k: np.load(rows[0][2])[k] * (2 if k == "emission_cube" else 1) for k in np.load(rows[0][2]).files})
    rows2 = sweep.tb_spectra(model, tmp_path, start_from_idx=2, **kw)
    assert len(rows2) == 4
    np.testing.assert_array_equal(np.load(rows2[0][2])["emission_cube"], 2 * first)
    np.testing.assert_array_equal(np.load(rows2[3][2])["emission_cube"], np.load(rows[3][2])["emission_cube"])
