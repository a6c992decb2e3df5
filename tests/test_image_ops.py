"""Image-plane post-processing: Gaussian beam and NaN patching.

CPU: the numpy restatements in oracle/ against golden outputs of the reference's own
patch_nan_emission_map (raytracingGRFF/util.py) and of scipy.ndimage.gaussian_filter, the call the
workflow makes (tests/golden/make_golden.py).  GPU: the CUDA kernels behind the drop-in
raytracinggrff_b200.util against both — the patched maps bit-exact, the beam to 1e-13 relative
(same summation order; the weights' exp/normalisation are evaluated by a different libm)."""
import numpy as np
import pytest

import cases


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


@pytest.mark.parametrize("name", cases.IMAGE_CASES)
def test_oracle_patch_nan_matches_reference(oracle, golden, name):
    out = oracle.patch_nan_emission_map(cases.image_case(name))
    assert _same(out, golden("image_ops")[f"patch_{name}"])


def test_oracle_patch_nan_semantics(oracle):
    a = np.array([[1.0, np.nan, 3.0], [np.nan, np.nan, np.nan], [7.0, np.nan, 9.0]])
    out = oracle.patch_nan_emission_map(a)
    # row-major, in place: (0,1) = mean(1, 3) [no finite pixel below it yet... (2,1) is NaN]; (1,0) = mean(1, 7)
    assert out[0, 1] == 2.0 and out[1, 0] == 4.0
    # (1,1): left is the just-patched (1,0) = 4, right none yet, down (0,1) = 2, up none -> 3
    assert out[1, 1] == 3.0
    assert np.isfinite(out).all()
    with pytest.raises(ValueError):
        oracle.patch_nan_emission_map(np.zeros(4))
    # nothing finite anywhere: left as it is
    assert np.isnan(oracle.patch_nan_emission_map(np.full((3, 4), np.nan))).all()


@pytest.mark.parametrize("sigma", cases.BEAM_SIGMAS)
def test_oracle_gaussian_filter_matches_scipy(oracle, golden, sigma):
    g = golden("image_ops")
    smooth = g["patch_sparse"]
    np.testing.assert_array_equal(oracle.gaussian_filter(smooth, sigma), g[f"beam_{sigma}"])


def test_oracle_gaussian_filter_properties(oracle):
    rng = np.random.default_rng(3)
    a = rng.random((30, 41))
    out = oracle.gaussian_filter(a, 3.0)
    # 'reflect' boundary conserves the total; a constant map is a fixed point
    assert abs(out.sum() - a.sum()) < 1e-9 * a.sum()
    np.testing.assert_allclose(oracle.gaussian_filter(np.full((9, 7), 2.5), 4.0), 2.5, rtol=1e-14)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.IMAGE_CASES)
def test_gpu_patch_nan_bit_exact(oracle, golden, name):
    from raytracinggrff_b200.util import patch_nan_emission_map
    a = cases.image_case(name)
    keep = a.copy()
    out = patch_nan_emission_map(a)
    assert _same(a, keep), "input must not be modified unless inplace"
    assert out.shape == a.shape and out.dtype == np.float64
    assert _same(out, golden("image_ops")[f"patch_{name}"])
    assert _same(out, oracle.patch_nan_emission_map(keep))
    b = keep.copy()
    assert patch_nan_emission_map(b, inplace=True) is b and _same(b, out)


@pytest.mark.gpu
def test_gpu_patch_nan_edge_cases(oracle):
    from raytracinggrff_b200.util import patch_nan_emission_map
    with pytest.raises(ValueError):
        patch_nan_emission_map(np.zeros(5))
    assert np.isnan(patch_nan_emission_map(np.full((4, 6), np.nan))).all()
    clean = np.arange(12.0).reshape(3, 4)
    np.testing.assert_array_equal(patch_nan_emission_map(clean), clean)
    one = np.array([[np.nan]])
    assert np.isnan(patch_nan_emission_map(one)).all()
    # a large random map, idempotence and agreement with the oracle
    rng = np.random.default_rng(5)
    big = rng.random((300, 257))
    big[rng.random(big.shape) < 0.2] = np.nan
    big[100:140, 50:120] = np.nan
    out = patch_nan_emission_map(big)
    assert _same(out, oracle.patch_nan_emission_map(big))
    assert np.isfinite(out).all()
    np.testing.assert_array_equal(patch_nan_emission_map(out), out)


@pytest.mark.gpu
@pytest.mark.parametrize("sigma", cases.BEAM_SIGMAS)
def test_gpu_gaussian_beam_matches_scipy(oracle, golden, sigma):
    from raytracinggrff_b200.util import gaussian_beam
    g = golden("image_ops")
    out = gaussian_beam(g["patch_sparse"], sigma)
    np.testing.assert_allclose(out, g[f"beam_{sigma}"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(out, oracle.gaussian_filter(g["patch_sparse"], sigma), rtol=1e-13, atol=0)


@pytest.mark.gpu
def test_gpu_gaussian_beam_cube_nan_and_workflow_wrappers(oracle, golden):
    from raytracinggrff_b200.util import apply_baseline_beam, convolve_beam, gaussian_beam
    g = golden("image_ops")
    # NaN pixels spread exactly as in scipy
    out = gaussian_beam(cases.image_case("sparse"), 2.5)
    assert np.array_equal(np.isnan(out), np.isnan(g["beam_nan_2.5"]))
    # (ny, nx, nf) cube: per-slice filtering
    rng = np.random.default_rng(9)
    cube = rng.random((40, 52, 3))
    out = gaussian_beam(cube, 1.7)
    for k in range(3):
        np.testing.assert_allclose(out[:, :, k], oracle.gaussian_filter(cube[:, :, k], 1.7), rtol=1e-13)
    # --consider-beam: sigma = fwhm / extent * N_pix (script/resample_with_ray_tracing.py:618-624)
    m = rng.random((64, 64))
    np.testing.assert_allclose(convolve_beam(m, 0.2, [-1.44, 1.44], 64),
                               oracle.gaussian_filter(m, 0.2 / 2.88 * 64), rtol=1e-13)
    # baseline beam (script/pub/compare_on_off_scaling_factor.py:51-69)
    xc = np.linspace(-1.44, 1.44, 64) * 6.957e8
    lam = 2.99792458e8 / 75e6
    sigma = lam / 3e3 * 1.495978707e11 / 6.957e8 / (2.88 / 63) / 2.355
    np.testing.assert_allclose(apply_baseline_beam(m, xc, xc, 75e6, 3.0), oracle.gaussian_filter(m, sigma), rtol=1e-13)
    np.testing.assert_array_equal(apply_baseline_beam(m, xc, xc, 75e6, 0.0), m)
