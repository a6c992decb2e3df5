"""Analytic known-answer checks for a GET_MW implementation (PARITY UNPINNED against the real
GRFF binary, which is absent; see oracle/oracle_grff.c).  Each check takes
``get_mw(Lparms, Rparms, Parms, T, DEM, DDM, RL) -> rc`` with the reference's array contract
(script/resample_with_ray_tracing.py:79-86, :489-509) so the same checks run against the CPU
oracle (tests/test_oracle_grff.py) and the CUDA library's PyGET_MW (tests/test_gpu_parity.py)."""
import numpy as np

E = 4.803204712570263e-10
ME = 9.1093837015e-28
C = 2.99792458e10
KB = 1.380649e-16
AU = 1.495978707e13
ZETA = (1 + 4 * 0.085) / (1 + 2 * 0.085)
KFF = 8 * E ** 6 / (3 * np.sqrt(2 * np.pi) * C * (ME * KB) ** 1.5)
AREA = 1.0e18


def parms(nz, dz, T, ne, B, theta=90.0, flag=5, smax=30):
    P = np.zeros((15, nz), dtype="double", order="F")
    P[0], P[1], P[2], P[3], P[4], P[6], P[7] = dz, T, ne, B, theta, flag, smax
    return P


def run(get_mw, P, f0, nf=1, step=0.0, area=AREA):
    L = np.array([P.shape[1], nf, 0, 0, 0], dtype="int32")
    R = np.array([area, f0, step], dtype="double")
    RL = np.zeros((7, nf), dtype="double", order="F")
    d = np.array(0, dtype="double")
    rc = get_mw(L, R, P, d, d, d, RL)
    assert rc == 0
    return RL


def tb_of(RL, area=AREA, pair=(5, 6)):
    """Brightness temperature with the LIBRARY constants (exact inverse of the flux conversion)."""
    nu = RL[0] * 1e9
    I = (RL[pair[0]] + RL[pair[1]]) * 1e-19 * AU * AU / area
    return I * C * C / (2 * KB * nu * nu)


def lnL(T, nu):
    return 18.2 + 1.5 * np.log(T) - np.log(nu) if T < 2e5 else 24.573 + np.log(T / nu)


def check_frequency_grid_and_codes(get_mw):
    P = parms(4, 1e8, 1e6, 1e8, 0.0)
    RL = run(get_mw, P, 450e6, nf=4, step=0.1)
    np.testing.assert_allclose(RL[0], 0.45 * 10 ** (0.1 * np.arange(4)), rtol=1e-14)
    d = np.array(0, dtype="double")
    RL1 = np.zeros((7, 1), order="F")
    assert get_mw(np.array([4, 0, 0, 0, 0], dtype="int32"), np.array([AREA, 1e9, 0.0]), P, d, d, d, RL1) == 1
    assert get_mw(np.array([4, 1, 3, 0, 0], dtype="int32"), np.array([AREA, 1e9, 0.0]), P, d, d, d, RL1) == 2
    # zero voxels: no emission
    RL0 = np.zeros((7, 1), order="F")
    assert get_mw(np.array([0, 1, 0, 0, 0], dtype="int32"), np.array([AREA, 1e9, 0.0]), P, d, d, d, RL0) == 0
    assert np.all(RL0[1:] == 0.0) and RL0[0, 0] == 1.0


def check_optically_thick_isothermal(get_mw):
    T, ne, nu = 1.0e6, 1.0e9, 1.0e9
    RL = run(get_mw, parms(100, 1e10, T, ne, 0.0), nu)
    v = E * E * ne / (np.pi * ME) / nu ** 2
    np.testing.assert_allclose(tb_of(RL)[0], T * (1 - v), rtol=1e-10)     # S = n^2 nu^2 k T / c^2 per mode
    assert abs(RL[5, 0] - RL[6, 0]) <= 1e-12 * RL[5, 0]                     # unpolarised for B = 0


def check_optically_thin_free_free(get_mw):
    T, ne, nu, dz, nz = 2.0e6, 1.0e8, 5.0e9, 1.0e7, 50
    RL = run(get_mw, parms(nz, dz, T, ne, 0.0), nu)
    v = E * E * ne / (np.pi * ME) / nu ** 2
    kap = KFF * ne ** 2 * ZETA * lnL(T, nu) / (np.sqrt(1 - v) * nu ** 2 * T ** 1.5)
    tau = kap * dz * nz
    assert tau < 1e-3
    np.testing.assert_allclose(tb_of(RL)[0], T * (1 - v) * (1 - np.exp(-tau)), rtol=1e-9)
    # low-temperature branch of the Coulomb logarithm
    T2 = 5.0e4
    RL2 = run(get_mw, parms(nz, dz, T2, ne, 0.0), nu)
    kap2 = KFF * ne ** 2 * ZETA * lnL(T2, nu) / (np.sqrt(1 - v) * nu ** 2 * T2 ** 1.5)
    np.testing.assert_allclose(tb_of(RL2)[0], T2 * (1 - v) * (1 - np.exp(-kap2 * dz * nz)), rtol=1e-9)


def check_polarisation_sign_and_magnitude(get_mw):
    T, ne, nu, dz, nz, B = 2.0e6, 1.0e8, 5.0e9, 1.0e7, 20, 50.0
    su = E * B / (2 * np.pi * ME * C) / nu
    for theta in (30.0, 150.0):
        RL = run(get_mw, parms(nz, dz, T, ne, B, theta), nu)
        vi = (RL[5, 0] - RL[6, 0]) / (RL[5, 0] + RL[6, 0])
        ct = np.cos(np.radians(theta))
        # thin, quasi-longitudinal: kappa_X,O ~ 1/(1 -+ sqrt(u)|cos|)^2, X = R for theta < 90, V=(L-R)/(L+R)
        np.testing.assert_allclose(vi, -2 * su * ct, rtol=2e-2)
    # theta = 90: F_O = 1, F_X = (u + (1-v)^2)/(1-v-u)^2 ; X is R (cos(90 deg) rounds to +6e-17)
    RL = run(get_mw, parms(nz, dz, T, ne, B, 90.0), nu)
    u, v = su ** 2, E * E * ne / (np.pi * ME) / nu ** 2
    fx = (u + (1 - v) ** 2) / (1 - v - u) ** 2
    nx2 = 1 - 2 * v * (1 - v) / (2 * (1 - v) - 2 * u)
    ratio = RL[6, 0] / RL[5, 0]          # R/L = X/O, thin: (kappa n^2)_X / (kappa n^2)_O
    np.testing.assert_allclose(ratio, fx * np.sqrt(nx2) / np.sqrt(1 - v), rtol=1e-6)
    # B -> 0: unpolarised
    RL0 = run(get_mw, parms(nz, dz, T, ne, 0.0, 30.0), nu)
    assert RL0[5, 0] == RL0[6, 0]


def check_cutoff_blocks_background(get_mw):
    nu = 100e6
    ne_crit = nu ** 2 * np.pi * ME / E ** 2
    P = parms(3, 1e9, 1e6, [0.5 * ne_crit, 2.0 * ne_crit, 0.1 * ne_crit], 0.0)
    RL = run(get_mw, P, nu)
    only_last = run(get_mw, parms(1, 1e9, 1e6, 0.1 * ne_crit, 0.0), nu)
    np.testing.assert_allclose(RL[1:], only_last[1:], rtol=1e-13)
    # empty voxels (n_e = 0 fill inside the Sun, zero padding) are transparent
    P2 = parms(3, [1e9, 0.0, 1e9], [1e6, 0.0, 1e6], [0.3 * ne_crit, 0.0, 0.3 * ne_crit], 0.0)
    P3 = parms(2, 1e9, 1e6, 0.3 * ne_crit, 0.0)
    np.testing.assert_allclose(run(get_mw, P2, nu)[1:], run(get_mw, P3, nu)[1:], rtol=1e-13)


def check_mode_coupling_limits(get_mw):
    # polarised thin emission from the far voxel (theta=60), then a quasi-transverse crossing into theta=120
    T, ne, dz, B = 2e6, 1e8, 1e7, 200.0
    P = parms(2, dz, T, [ne, 1e-3], B, [60.0, 120.0])
    for nu, expect in ((5e10, "strong"), (1.5e9, "weak")):
        RL = run(get_mw, P, nu)
        assert RL[1, 0] > 0 and RL[2, 0] > 0
        # weak coupling swaps L and R relative to strong coupling
        np.testing.assert_allclose([RL[1, 0], RL[2, 0]], [RL[4, 0], RL[3, 0]], rtol=1e-6)
        dzm = dz
        g = np.radians(60.0) / dzm
        d = E ** 5 / (32 * np.pi ** 2 * ME ** 4 * C ** 4) * (0.5 * (ne + 1e-3)) * B ** 3 / (nu ** 4 * g)
        Q = np.exp(-d)
        np.testing.assert_allclose(RL[5, 0], Q * RL[3, 0] + (1 - Q) * RL[4, 0], rtol=1e-9)
        np.testing.assert_allclose(RL[6, 0], Q * RL[4, 0] + (1 - Q) * RL[3, 0], rtol=1e-9)
        if expect == "strong":
            assert Q > 0.999
        else:
            assert Q < 1e-3


def check_gyroresonance_layer(get_mw):
    # B falls through the s=3 layer of nu between two voxels; FF switched off (flag bit 1)
    nu, T, ne, dz, theta = 5.0e9, 3.0e6, 2.0e9, 2.0e8, 40.0
    Bres = nu * 2 * np.pi * ME * C / (3 * E)
    Bp, Bk = 1.08 * Bres, 0.95 * Bres          # contains s=3 only (s=2 needs 1.5 Bres, s=4 0.75 Bres)
    P = parms(2, dz, T, ne, [Bp, Bk], theta, flag=2 + 4)
    RL = run(get_mw, P, nu)
    th = np.radians(theta)
    ct, st = np.cos(th), np.sin(th)
    u, v = 1.0 / 9.0, E * E * ne / (np.pi * ME) / nu ** 2
    LB = Bres * dz / abs(Bk - Bp)
    beta2 = KB * T / (ME * C * C)
    I = {}
    for sg in (-1, 1):
        D = u * u * st ** 4 + 4 * u * (1 - v) ** 2 * ct ** 2
        sD = sg * np.sqrt(D)
        n2 = 1 - 2 * v * (1 - v) / (2 * (1 - v) - u * st ** 2 + sD)
        Ts = 2 * np.sqrt(u) * (1 - v) * ct / (u * st ** 2 - sD)
        Ls = (v * np.sqrt(u) * st + Ts * u * v * st * ct) / (1 - u - v + u * v * ct ** 2)
        s = 3
        tau = (np.pi * E * E * ne * LB / (ME * C * nu) * s ** (2 * s) / (2 ** (s - 1) * 6.0)
               * (beta2 * st * st) ** (s - 1) * n2 ** (s - 1.5) * (Ts * ct + Ls * st + 1) ** 2 / (1 + Ts * Ts))
        I[sg] = n2 * nu * nu * KB * T / (C * C) * (1 - np.exp(-tau))
        assert tau > 1e-4
    to_sfu = AREA / AU ** 2 / 1e-19
    np.testing.assert_allclose(RL[6, 0], I[-1] * to_sfu, rtol=1e-9)      # X is R for theta < 90
    np.testing.assert_allclose(RL[5, 0], I[+1] * to_sfu, rtol=1e-9)
    assert RL[6, 0] > 2 * RL[5, 0]                                        # X more opaque than O
    # GR switched off (flag bit 0): nothing is emitted
    RL_off = run(get_mw, parms(2, dz, T, ne, [Bp, Bk], theta, flag=1 + 2 + 4), nu)
    assert np.all(RL_off[1:] == 0.0)
    # s_max below the harmonic: no layer
    RL_s2 = run(get_mw, parms(2, dz, T, ne, [Bp, Bk], theta, flag=2 + 4, smax=2), nu)
    assert np.all(RL_s2[1:] == 0.0)


def check_s_input_scales_the_source(get_mw):
    """Parms[14] > 0 = the voxel's own source area S*area (--s-input-on, script/...:501).  Defined in
    oracle/oracle_grff.c: the voxel's source term is multiplied by Parms[14]/Rparms[0], absorption
    unchanged.  Optically thin: emission scales by S; optically thick uniform S: T_b -> S*T; a
    foreground voxel with S = 1 in front of a thick S = 3 slab absorbs like any other."""
    T, ne, nu = 1.0e6, 2.0e8, 3.0e9
    thin = parms(8, 1e7, T, ne, 0.0)
    base = run(get_mw, thin, nu)
    for S in (0.5, 1.0, 2.5):
        P = thin.copy(order="F")
        P[14] = S * AREA
        RL = run(get_mw, P, nu)
        np.testing.assert_allclose(RL[5:], S * base[5:], rtol=2e-6)      # thin: tau ~ 1e-5, (1 - S tau/2 ...) ~ 1
    # Parms[14] = 0 (the reference's default packing) and Parms[14] = area are the same map
    P = thin.copy(order="F")
    P[14] = AREA
    np.testing.assert_allclose(run(get_mw, P, nu)[5:], base[5:], rtol=1e-14)
    thick = parms(64, 1e9, T, 3e9, 0.0)
    thick[14] = 3.0 * AREA
    np.testing.assert_allclose(tb_of(run(get_mw, thick, 1.0e9))[0], 3.0 * T * (1 - 8.06163860e7 * 3e9 / 1e18), rtol=2e-3)
    # per-voxel factors: S ramps along a thin LOS -> emission = sum S_k * e_k
    P = thin.copy(order="F")
    Sk = np.linspace(0.2, 1.8, 8)
    P[14] = Sk * AREA
    np.testing.assert_allclose(run(get_mw, P, nu)[5:], Sk.mean() * base[5:], rtol=1e-5)


def check_literature_numbers(get_mw):
    """Anchors OUTSIDE this repo's own restatement: the closed-form approximations of the literature with their
    PUBLISHED numerical constants (not the constants of oracle_grff.c), in the regimes where they hold.
      * free-free: Dulk (1985, ARA&A 23, 169) eqs. 20-21: kappa = 9.78e-3 n_e^2 / (nu^2 T^1.5) x (18.2 + ln T^1.5 - ln nu)
        for T < 2e5 K, x (24.5 + ln T - ln nu) above, for a hydrogen plasma: the implementation carries sum(Z^2 n_i)/n_e
        for H + He (1.145) on top and 24.573 for 24.5.
      * gyroresonance: White & Kundu (1997, Sol. Phys. 174, 31) eq. 1 / Dulk (1985) eq. 36:
        tau_{X,O} = 0.0133 n_e L_B / nu x (s^2 / s!) (s^2 sin^2 theta / (2 mu))^(s-1) (1 -+ sigma |cos theta|)^2,
        mu = m c^2 / k T = 5.93e9 / T, valid for quasi-circular modes and nu >> nu_p.
    Agreement is expected to the accuracy of those approximations (a percent or a few), not to rounding."""
    to_sfu = AREA / AU ** 2 / 1e-19
    # --- free-free, both Coulomb-logarithm branches, optically thin, nu >> nu_p
    for T in (5.0e4, 2.0e6):
        ne, nu, dz = 1.0e8, 5.0e9, 1.0e8
        RL = run(get_mw, parms(1, dz, T, ne, 0.0), nu)
        src = nu * nu * KB * T / (C * C)                                  # n ~ 1
        tau = (RL[5, 0] + RL[6, 0]) / to_sfu / (2 * src)
        coul = 18.2 + 1.5 * np.log(T) - np.log(nu) if T < 2e5 else 24.5 + np.log(T) - np.log(nu)
        tau_dulk = 9.78e-3 * ne * ne / (nu * nu * T ** 1.5) * coul * dz
        assert tau < 1e-3
        np.testing.assert_allclose(tau / ZETA, tau_dulk, rtol=6e-3)      # He correction apart: within 0.6 %
    # --- gyroresonance, s = 2 and 3, theta = 40 deg (quasi-circular: u sin^4 / (4 cos^2) << 1), FF off, layers thin
    for s, fact, ne in ((2, 2.0, 2.0e4), (3, 6.0, 5.0e6)):
        nu, T, dz, theta = 5.0e9, 3.0e6, 2.0e8, 40.0
        Bres = nu / (2.80e6 * s)                                            # nu_B = 2.80 MHz per gauss
        lo = {2: 0.80, 3: 0.95}[s]                                          # keeps the neighbouring harmonics out
        Bp, Bk = 1.08 * Bres, lo * Bres
        RL = run(get_mw, parms(2, dz, T, ne, [Bp, Bk], theta, flag=2 + 4), nu)
        th = np.radians(theta)
        LB = Bres * dz / abs(Bk - Bp)
        mu = 5.93e9 / T
        src = nu * nu * KB * T / (C * C)
        tau, tau_wk = {}, {}
        for row, sigma in ((6, -1), (5, +1)):                               # X (sigma = -1) is R for theta < 90
            tau[sigma] = -np.log1p(-RL[row, 0] / to_sfu / src)
            tau_wk[sigma] = (0.0133 * ne * LB / nu * (s * s / fact) * (s * s * np.sin(th) ** 2 / (2 * mu)) ** (s - 1)
                             * (1 - sigma * abs(np.cos(th))) ** 2)
        assert 0.02 < tau[-1] < 0.3
        np.testing.assert_allclose(tau[-1], tau_wk[-1], rtol=0.02)          # X mode: the published number to 2 %
        # O mode: the (1 - |cos|)^2 factor of the quasi-circular approximation is known to overestimate it; the exact
        # polarisation coefficients used here give a few times less — bounded, and far below the X mode
        assert tau_wk[+1] / 6 < tau[+1] < tau_wk[+1] and tau[+1] < 0.05 * tau[-1]


ALL_CHECKS = [check_literature_numbers, check_s_input_scales_the_source, check_frequency_grid_and_codes, check_optically_thick_isothermal, check_optically_thin_free_free,
              check_polarisation_sign_and_magnitude, check_cutoff_blocks_background, check_mode_coupling_limits,
              check_gyroresonance_layer]


def random_los_batch(rng, npix, nz, with_b=True, flag=5, theta90=True):
    """Random but physical batched input in the fastGRFF layout (script/...:404-446) with ragged
    valid counts and zero padding."""
    Parms_M = np.zeros((15, nz, npix), dtype=np.float64, order="F")
    Parms_M[4] = 90.0
    Parms_M[6] = flag
    Parms_M[7] = 30
    for p in range(npix):
        cnt = int(rng.integers(0, nz + 1))
        Parms_M[0, :cnt, p] = 10 ** rng.uniform(7.5, 9.5, cnt)
        Parms_M[1, :cnt, p] = 10 ** rng.uniform(4.5, 6.8, cnt)
        Parms_M[2, :cnt, p] = 10 ** rng.uniform(6.0, 9.3, cnt)
        if with_b:
            Parms_M[3, :cnt, p] = 10 ** rng.uniform(-1.0, 2.8, cnt)
        if not theta90:
            Parms_M[4, :cnt, p] = np.clip(np.cumsum(rng.normal(0, 12, cnt)) + rng.uniform(20, 160), 1.0, 179.0)
    return Parms_M
