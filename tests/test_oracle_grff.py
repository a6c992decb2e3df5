"""CPU: analytic known-answer tests of the GRFF restatement (oracle/oracle_grff.c).
PARITY UNPINNED — the real GRFF binary is absent; these pin the restatement to the published
formulae and to the reference's array contract."""
import numpy as np
import pytest

import grff_checks


@pytest.mark.parametrize("check", grff_checks.ALL_CHECKS, ids=lambda f: f.__name__)
def test_oracle_grff_analytic(oracle, check):
    check(oracle.get_mw)


def test_oracle_batched_equals_single(oracle):
    rng = np.random.default_rng(3)
    npix, nz, nf = 17, 40, 3
    P = grff_checks.random_los_batch(rng, npix, nz, theta90=False, flag=4)
    L = np.array([npix, nz, nf, 1, 0, 0], dtype=np.int32)
    R = np.zeros((3, npix), order="F")
    R[0], R[1], R[2] = 2.5e17, 3e8, 0.2
    RL_M = np.zeros((7, nf, npix), order="F")
    status = oracle.get_mw_slice(L, R, P, None, None, None, RL_M)
    assert np.all(status == 0)
    for p in range(npix):
        RL = np.zeros((7, nf), order="F")
        cnt = int(np.count_nonzero(P[0, :, p] > 0))
        assert oracle.get_mw(np.array([cnt, nf, 0, 0, 0], dtype=np.int32), R[:, p].copy(),
                             np.asfortranarray(P[:, :cnt, p]), None, None, None, RL) == 0
        np.testing.assert_array_equal(RL, RL_M[:, :, p])   # zero padding is transparent


def test_workflow_conversion_constants(oracle):
    """T_b through the workflow's own constants (script/resample_with_ray_tracing.py:91-94, :513-520)
    differs from the library-constant inverse by the known ratio only."""
    P = grff_checks.parms(100, 1e10, 1e6, 1e9, 0.0)
    RL = grff_checks.run(oracle.get_mw, P, 1e9)
    conv = (1e-19 * 2.998e10 ** 2 / (2 * 1.38065e-16 * 1e18) / grff_checks.AREA) * 1.49599e13 ** 2
    tb_workflow = (RL[5, 0] + RL[6, 0]) * conv
    ratio = (2.998e10 / grff_checks.C) ** 2 * (grff_checks.KB / 1.38065e-16) * (1.49599e13 / grff_checks.AU) ** 2
    np.testing.assert_allclose(tb_workflow, grff_checks.tb_of(RL)[0] * ratio, rtol=1e-12)
