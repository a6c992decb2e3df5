"""CPU: the float32 voxel evaluation of the per-ray kernels (raytracinggrff_b200/csrc/grff_fast.cuh) — the very
source the CUDA kernels compile, built for the host with g++ — against a float64 numpy restatement of the same
published formulas and against the oracle's GET_MW on single-voxel lines of sight.

Where the fast path reports `ok` its slab operator (a = e^-tau, b = S (1 - e^-tau) per mode) must agree with
float64 to a few 1e-6 — two orders below the 1e-4 tolerance on T_b; where it does not (near a mode cut-off or
the gyro-resonance) the kernels evaluate that voxel in FP64, and the test checks that those regions — and only
a small share of an adversarial random sample — are the ones it declines."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
K_NUB = 2.79924898723330395e+06
K_NUP2 = 8.06163860001142621e+07
K_FF = 9.76981314722991795e-03
K_ZETA = 1.14529914529914545
K_KBC2 = 1.53617918724037216e-37


@pytest.fixture(scope="module")
def fast():
    src = HERE / "native" / "grff_fast_host.cpp"
    so = HERE / "native" / "libgrff_fast_host.so"
    hdr = HERE.parent / "raytracinggrff_b200" / "csrc" / "grff_fast.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    fp = ctypes.POINTER(ctypes.c_float)
    lib.grff_fast_eval.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_int, fp, ctypes.POINTER(ctypes.c_uint8)]

    def run(nu, vox, ff_on=True):
        cv = K_NUP2 / nu ** 2
        hi = np.float32(cv)
        fc = np.array([K_NUB / nu, hi, cv - float(hi), 18.2 - np.log(nu), 24.573 - np.log(nu), K_FF * K_ZETA / nu ** 2,
                       nu ** 2 * K_KBC2], dtype=np.float32)
        vox = np.ascontiguousarray(vox, dtype=np.float32)
        out = np.zeros((len(vox), 4), np.float32)
        ok = np.zeros(len(vox), np.uint8)
        lib.grff_fast_eval(fc.ctypes.data_as(fp), vox.ctypes.data_as(fp), len(vox), int(ff_on), out.ctypes.data_as(fp),
                           ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return out, ok.astype(bool)
    return run


def reference(nu, dz, T, ne, B, cth, sth):
    """float64 restatement (per mode: n^2, free-free kappa, Kirchhoff source, slab), L/R assignment by sign(cos theta)."""
    dz, T, ne, B, cth, sth = (np.asarray(a, dtype=np.float64) for a in (dz, T, ne, B, cth, sth))
    u, v = (K_NUB * B / nu) ** 2, K_NUP2 * ne / nu ** 2
    omv, s2, c2 = 1 - v, sth * sth, cth * cth
    lnL = np.where(T < 2e5, 18.2 + 1.5 * np.log(T) - np.log(nu), 24.573 + np.log(T / nu))
    pref = K_FF * ne * ne * K_ZETA * lnL / (nu * nu * T * np.sqrt(T))
    srcb = nu * nu * K_KBC2 * T
    res = []
    with np.errstate(all="ignore"):
        sD = np.sqrt(u * u * s2 * s2 + 4 * u * omv * omv * c2)
        for sg in (-1, 1):
            den = 2 * omv - u * s2 + sg * sD
            n2 = np.where(u > 0, 1 - 2 * v * omv / den, omv)
            F = np.where(u > 0, 2 * (sg * sD * (u * s2 + 2 * omv * omv) - u * u * s2 * s2) / (sg * sD * den * den), 1.0)
            on = np.where(u > 0, ~((u >= 1) | (v >= 1 - np.sqrt(u))), v < 1) if sg < 0 else (v < 1)
            on = on & (n2 > 0) & np.isfinite(n2) & np.isfinite(F)
            kap = pref * F / np.sqrt(n2)
            kap = np.where((kap > 0) & np.isfinite(kap), kap, 0.0)
            tau = kap * dz
            res.append((np.where(on, np.exp(-tau), 0.0), np.where(on, n2 * srcb * (-np.expm1(-tau)), 0.0), n2, on))
    (aX, bX, n2X, onX), (aO, bO, n2O, _) = res
    xr = cth >= 0
    return np.where(xr, aO, aX), np.where(xr, aX, aO), np.where(xr, bO, bX), np.where(xr, bX, bO), n2X, n2O, onX, omv, srcb


def random_voxels(rng, nu, n):
    v = 10 ** rng.uniform(-5, 0.05, n)
    ne = v * nu ** 2 / K_NUP2
    B = 10 ** rng.uniform(-3, 3.5, n)
    B[::17] = 0
    th = rng.uniform(0, np.pi, n)
    cth, sth = np.cos(th), np.sin(th)
    cth[::5], sth[::5] = 6.123233995736766e-17, 1.0        # the reference's theta = 90 deg
    cth[1::31], sth[1::31] = 1.0, 0.0
    T = 10 ** rng.uniform(4, 6.6, n)
    dz = 10 ** rng.uniform(6, 10.5, n)
    return np.stack([dz, T, ne, B, cth, sth, np.ones(n)], axis=1).astype(np.float32)


@pytest.mark.parametrize("nu", [20e6, 75e6, 300e6, 1.5e9])
def test_fast_voxel_matches_float64_where_it_accepts(fast, nu):
    rng = np.random.default_rng(int(nu) % 1000)
    vox = random_voxels(rng, nu, 200000)
    out, ok = fast(nu, vox)
    aL, aR, bL, bR, n2X, n2O, onX, omv, srcb = reference(nu, *vox[:, :6].T)
    assert ok.mean() > 0.97                                    # even on this adversarial sample
    da = np.maximum(np.abs(out[:, 0] - aL), np.abs(out[:, 1] - aR))[ok]
    db = (np.maximum(np.abs(out[:, 2] - bL), np.abs(out[:, 3] - bR)) / srcb)[ok]
    assert da.max() < 5e-6 and db.max() < 5e-6, (da.max(), db.max())
    assert np.quantile(da, 0.999) < 5e-7 and np.quantile(db, 0.999) < 5e-7
    # what it declines: close to the plasma cut-off, a mode's own cut-off, or the gyro-resonance
    declined = ~ok
    near = (omv < 0.03) | (n2O < 0.03) | (onX & (n2X < 0.03)) | (np.abs(np.sqrt((K_NUB * vox[:, 3] / nu) ** 2) - 1) < 0.15) | ~np.isfinite(n2X)
    assert (declined & ~near).mean() < 2e-3
    # and it never accepts a voxel whose mode is within the guard of its cut-off
    assert not (ok & (omv < 0.019)).any()


def test_fast_voxel_against_oracle_get_mw_single_voxel(fast, oracle):
    """One-voxel lines of sight through the oracle's GET_MW: RL[5], RL[6] are then b_L, b_R in sfu."""
    rng = np.random.default_rng(3)
    nu, area = 150e6, 3.0e17
    vox = random_voxels(rng, nu, 3000)
    vox[:, 4], vox[:, 5] = np.cos(np.deg2rad(60.0)), np.sin(np.deg2rad(60.0))
    out, ok = fast(nu, vox)
    P = np.zeros((15, 1, len(vox)), order="F")
    P[0, 0], P[1, 0], P[2, 0], P[3, 0], P[4, 0], P[6, 0], P[7, 0] = vox[:, 0], vox[:, 1], vox[:, 2], vox[:, 3], 60.0, 5, 30
    L = np.array([len(vox), 1, 1, 1, 0, 0], dtype=np.int32)
    R = np.zeros((3, len(vox)), order="F")
    R[0], R[1] = area, nu
    RL = np.zeros((7, 1, len(vox)), order="F")
    oracle.get_mw_slice(L, R, P, None, None, None, RL)
    to_sfu = area / 1.495978707e13 ** 2 / 1e-19
    src = nu * nu * K_KBC2 * vox[:, 1].astype(np.float64) * to_sfu
    # theta = 60 deg as float32 cos/sin vs the oracle's cos(60 deg) in double: 3e-8 apart
    assert np.abs(out[ok, 2] * to_sfu - RL[5, 0][ok]).max() / src.max() < 5e-6
    assert (np.abs(out[:, 2] * to_sfu - RL[5, 0]) / src)[ok].max() < 5e-6
    assert (np.abs(out[:, 3] * to_sfu - RL[6, 0]) / src)[ok].max() < 5e-6


def test_fast_voxel_edge_values(fast):
    nu = 100e6
    base = np.array([1e8, 1e6, 1e7, 2.0, 0.5, np.sqrt(0.75), 1.0], dtype=np.float32)
    vox = np.tile(base, (6, 1))
    vox[1, 3] = 0.0                     # B = 0: unpolarised
    vox[2, 2] = 1.3e8                   # v > 1: beyond the plasma cut-off -> declined (FP64 path says evanescent)
    vox[3, 3] = 35.7                    # nu_B = nu: X mode cut off, O propagates
    vox[4, 0] = 1e14                    # optically thick: a -> 0, b -> n^2 S
    vox[5, 6] = 2.5                     # S input scales the source only
    out, ok = fast(nu, vox)
    assert ok.tolist() == [True, True, False, True, True, True]
    assert out[1, 0] == out[1, 1] and out[1, 2] == out[1, 3]
    assert out[3, 1] == 0.0 and out[3, 3] == 0.0 and 0 < out[3, 0] < 1          # cos > 0: X is R
    assert out[4, 0] < 1e-30 and out[4, 1] < 1e-30
    np.testing.assert_allclose(out[5, 2:], 2.5 * out[0, 2:], rtol=1e-6)
    np.testing.assert_array_equal(out[5, :2], out[0, :2])
