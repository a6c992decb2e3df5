/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_raytrace.c header).
 *
 * CPU restatement (plain C, float64) of the GRFF "GET_MW" computation the reference
 * calls through ctypes:
 *   /root/reference/script/resample_with_ray_tracing.py:79-86   (PyGET_MW binding)
 *   /root/reference/script/resample_with_ray_tracing.py:489-520 (Parms layout, RL use)
 *   /root/reference/script/resample_with_ray_tracing.py:404-446 (fastGRFF batched layout)
 *   /root/reference/script/synthetic_FF_map_single_thread.py:189-220 (straight-LOS twin)
 *
 * PARITY UNPINNED.  The arithmetic behind PyGET_MW lives in a third-party dependency
 * that is NOT in /root/reference: GRFF_DEM_Transfer.so (kuznetsov-radio/GRFF, C++) and
 * its GPU sibling fastGRFF; the reference pins no version, tag or hash for either
 * (only a path, script/resample_with_ray_tracing.py:88-89; README.md:9,16) and holds no
 * test or golden vector at this boundary.  This file therefore restates the PUBLISHED
 * algorithm (Fleishman, Kuznetsov & Landi 2021, ApJ 914, 52; Fleishman & Kuznetsov
 * 2010, ApJ 721, 1127 App. A; Dulk 1985 ARA&A 23, 169; Zheleznyakov 1970) and anchors on
 * the reference's call-site contract (array layouts, flags, units, which RL rows are
 * read) plus analytic limits (tests/test_oracle_grff.py).  Differences to be expected
 * against the real GRFF binary: classical Coulomb logarithm instead of tabulated Gaunt
 * factors / abundance-dependent zeta(T); no DEM/DDM; no neutral (H, He) opacity.
 *
 * Model (per frequency nu, per voxel k = 0..Nz-1, voxel 0 farthest from the observer;
 * the emerging intensity is the one after voxel Nz-1):
 *   u=(nu_B/nu)^2, v=(nu_p/nu)^2, D=u^2 sin^4 th + 4u(1-v)^2 cos^2 th,
 *   n_s^2 = 1 - 2v(1-v)/(2(1-v) - u sin^2 th + s sqrt D),  s=-1 (X), +1 (O);
 *   free-free  kappa_s = K n_e^2 zeta lnL F_s / (n_s nu^2 T^1.5),
 *              K = 8 e^6 / (3 sqrt(2 pi) c (m k)^1.5), zeta=(1+4A)/(1+2A), A=He/H=0.085,
 *              lnL = 18.2+1.5 lnT-ln nu (T<2e5) | 24.573+ln(T/nu),
 *              F_s = 2 (s sqrtD [u sin^2 th+2(1-v)^2] - u^2 sin^4 th)
 *                      / (s sqrtD [2(1-v) - u sin^2 th + s sqrtD]^2)      (=1 for B=0),
 *              source S_s = n_s^2 nu^2 k T / c^2 (Kirchhoff per mode);
 *   slab:      I <- I e^-tau + S (1-e^-tau), tau = kappa dz;
 *   cutoff:    a mode with n^2<=0 (or X above its cutoff v >= 1-sqrt u, or u>=1) is
 *              evanescent in that voxel: I_s <- 0, no emission;
 *   empty voxel: dz<=0, T<=0, n_e<=0, B<0 or non-finite -> transparent; no layer is placed
 *              between an empty voxel and its neighbours;
 *   gyroresonance layers (Parms[6] bit0 clear) between consecutive voxels where
 *              s nu_B crosses nu, s=2..s_max: B, n_e, T, theta linear between voxel
 *              centres, L_B = B_res dz_mid/|dB|,
 *              tau_s = (pi e^2 n_e L_B/(m c nu)) s^2s/(2^(s-1) s!) (beta^2 sin^2 th)^(s-1)
 *                      n_s^(2s-3) (T_s cos th + L_s sin th + 1)^2/(1+T_s^2),
 *              T_s = 2 sqrt u (1-v) cos th/(u sin^2 th - s sqrtD),
 *              L_s = (v sqrt u sin th + T_s u v sin th cos th)/(1-u-v+u v cos^2 th);
 *   polarisation: X is R where cos th >= 0, L otherwise; where cos th changes sign
 *              between voxels (quasi-transverse layer): weak coupling swaps L,R; strong
 *              leaves them; exact mixes with Q=exp(-d),
 *              d = e^5/(32 pi^2 m^4 c^4) n_e B^3/(nu^4 |dtheta/dz|);
 *   S input (Parms[14] > 0, the reference's --s-input-on, script/resample_with_ray_tracing.py:501):
 *              the slot holds the voxel's source area S_k * area [cm^2], the cross-section of the
 *              ray pencil there.  The reference hands it to a private GRFF build whose use of it is
 *              not published; DEFINED HERE as: the voxel's emission into the pixel's flux scales
 *              with its own area, i.e. its source term is multiplied by Parms[14] / Rparms[0]
 *              (absorption unchanged; a gyroresonance layer between two voxels takes the factor
 *              interpolated like its other parameters).  Parms[14] <= 0 (the default packing) leaves the factor 1;
 *   output:    RL[0]=nu/1e9, RL[1,2]=L,R weak, RL[3,4]=strong, RL[5,6]=exact, in sfu for
 *              source area Rparms[0] seen from 1 au.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define Q_EL 4.803204712570263e-10
#define M_EL 9.1093837015e-28
#define C_L 2.99792458e10
#define K_B 1.380649e-16
#define AU_CM 1.495978707e13
#define SFU 1e-19
#define HE_ABUND 0.085

typedef struct {
    double dz, T, ne, B, th, cth, sth;
    double scale;                 /* source-term factor: Parms[14] / area when Parms[14] > 0, else 1 */
    int gr_on, ff_on, smax;
} voxel_t;

typedef struct { double n2, kap, src; int prop; } mode_t;

/* refractive index, FF opacity and source function of mode sg at (ne,B,T,theta). */
static void mode_eval(double nu, double ne, double B, double T, double cth, double sth,
                      int sg, int ff_on, mode_t *m, double *Tpol, double *Lpol)
{
    const double nup2 = Q_EL * Q_EL * ne / (M_PI * M_EL);
    const double nuB = Q_EL * B / (2.0 * M_PI * M_EL * C_L);
    const double u = (nuB / nu) * (nuB / nu), v = nup2 / (nu * nu);
    const double s2 = sth * sth, c2 = cth * cth;
    m->prop = 0; m->kap = 0.0; m->src = 0.0; m->n2 = 0.0;
    if (Tpol) { *Tpol = 0.0; *Lpol = 0.0; }
    double F = 1.0, n2;
    if (u > 0.0) {
        if (u >= 1.0 && sg < 0) return;                    /* nu <= nu_B: no escaping X mode */
        if (sg < 0 && v >= 1.0 - sqrt(u)) return;          /* above the X-mode cutoff */
        if (sg > 0 && v >= 1.0) return;                    /* above the O-mode cutoff */
        const double D = u * u * s2 * s2 + 4.0 * u * (1.0 - v) * (1.0 - v) * c2;
        const double sD = sg * sqrt(D);
        const double den = 2.0 * (1.0 - v) - u * s2 + sD;
        n2 = 1.0 - 2.0 * v * (1.0 - v) / den;
        F = 2.0 * (sD * (u * s2 + 2.0 * (1.0 - v) * (1.0 - v)) - u * u * s2 * s2) / (sD * den * den);
        if (Tpol) {
            const double Ts = 2.0 * sqrt(u) * (1.0 - v) * cth / (u * s2 - sD);
            *Tpol = Ts;
            *Lpol = (v * sqrt(u) * sth + Ts * u * v * sth * cth) / (1.0 - u - v + u * v * c2);
        }
    } else {
        if (v >= 1.0) return;
        n2 = 1.0 - v;
    }
    if (!(n2 > 0.0) || !isfinite(n2) || !isfinite(F)) return;
    m->prop = 1;
    m->n2 = n2;
    m->src = n2 * nu * nu * K_B * T / (C_L * C_L);
    if (ff_on && ne > 0.0) {
        const double K = 8.0 * pow(Q_EL, 6) / (3.0 * sqrt(2.0 * M_PI) * C_L * pow(M_EL * K_B, 1.5));
        const double zeta = (1.0 + 4.0 * HE_ABUND) / (1.0 + 2.0 * HE_ABUND);
        const double lnL = (T < 2e5) ? 18.2 + 1.5 * log(T) - log(nu) : 24.573 + log(T / nu);
        double kap = K * ne * ne * zeta * lnL * F / (sqrt(n2) * nu * nu * T * sqrt(T));
        if (!(kap > 0.0) || !isfinite(kap)) kap = 0.0;
        m->kap = kap;
    }
}

static inline void slab(double *I, double tau, double src)
{
    if (tau > 0.0) {
        const double em = -expm1(-tau);       /* 1 - e^-tau */
        *I = *I * (1.0 - em) + src * em;
    }
}

/* apply mode (a = e^-tau part expressed through tau, src) to the L or R slot of all
 * three coupling variants. I6 = {Lw,Rw,Ls,Rs,Le,Re}. */
static inline void apply_mode(double *I6, int to_R, int prop, double tau, double src)
{
    for (int q = 0; q < 3; ++q) {
        double *I = &I6[2 * q + (to_R ? 1 : 0)];
        if (!prop) *I = 0.0; else slab(I, tau, src);
    }
}

static void qt_layer(double *I6, double nu, const voxel_t *p, const voxel_t *k)
{
    const double dzm = 0.5 * (p->dz + k->dz);
    const double g = fabs(k->th - p->th) / dzm;
    const double nav = 0.5 * (p->ne + k->ne), Bav = 0.5 * (p->B + k->B);
    const double cq = pow(Q_EL, 5) / (32.0 * M_PI * M_PI * pow(M_EL, 4) * pow(C_L, 4));
    const double d = cq * nav * Bav * Bav * Bav / (nu * nu * nu * nu * g);
    const double Q = exp(-d);
    double t = I6[0]; I6[0] = I6[1]; I6[1] = t;          /* weak: swap */
    const double Le = I6[4], Re = I6[5];                 /* strong: unchanged */
    I6[4] = Q * Le + (1.0 - Q) * Re;
    I6[5] = Q * Re + (1.0 - Q) * Le;
}

static void gr_layer(double *I6, double nu, int s, double ne, double T, double th, double LB, double scale)
{
    const double cth = cos(th), sth = sin(th);
    const double Bres = nu * 2.0 * M_PI * M_EL * C_L / (s * Q_EL);
    const int x_to_R = cth >= 0.0;
    for (int sg = -1; sg <= 1; sg += 2) {
        mode_t m; double Ts, Ls;
        mode_eval(nu, ne, Bres, T, cth, sth, sg, 0, &m, &Ts, &Ls);
        const int to_R = (sg < 0) ? x_to_R : !x_to_R;
        if (!m.prop) { apply_mode(I6, to_R, 0, 0.0, 0.0); continue; }
        const double beta2 = K_B * T / (M_EL * C_L * C_L);
        const double lg = 2.0 * s * log((double)s) - (s - 1) * log(2.0) - lgamma(s + 1.0) +
                          (s - 1) * log(beta2 * sth * sth) + (s - 1.5) * log(m.n2);
        const double pol = Ts * cth + Ls * sth + 1.0;
        double tau = M_PI * Q_EL * Q_EL * ne * LB / (M_EL * C_L * nu) * exp(lg) * pol * pol / (1.0 + Ts * Ts);
        if (!(tau > 0.0) || !isfinite(tau)) tau = 0.0;
        apply_mode(I6, to_R, 1, tau, m.src * scale);
    }
}

/* events between voxel centres p -> k, in path order. */
static void between(double *I6, double nu, const voxel_t *p, const voxel_t *k)
{
    const int qt = (p->cth * k->cth < 0.0);
    const double tqt = qt ? (0.5 * M_PI - p->th) / (k->th - p->th) : 2.0;
    int qt_done = !qt;
    if (p->gr_on && k->gr_on && p->B != k->B) {
        const int smax = p->smax < k->smax ? p->smax : k->smax;
        const double dzm = 0.5 * (p->dz + k->dz);
        const int up = k->B > p->B;       /* B rising: high harmonics (small B_res) come first */
        for (int q = 2; q <= smax; ++q) {
            const int s = up ? (smax + 2 - q) : q;
            const double Bres = nu * 2.0 * M_PI * M_EL * C_L / (s * Q_EL);
            if ((p->B - Bres) * (k->B - Bres) >= 0.0) continue;
            const double t = (Bres - p->B) / (k->B - p->B);
            if (!qt_done && t >= tqt) { qt_layer(I6, nu, p, k); qt_done = 1; }
            const double ne = p->ne + t * (k->ne - p->ne), T = p->T + t * (k->T - p->T);
            const double th = p->th + t * (k->th - p->th);
            const double LB = Bres * dzm / fabs(k->B - p->B);
            gr_layer(I6, nu, s, ne, T, th, LB, p->scale + t * (k->scale - p->scale));
        }
    }
    if (!qt_done) qt_layer(I6, nu, p, k);
}

/*
 * Same contract as GRFF's PyGET_MW (script/resample_with_ray_tracing.py:79-86, :502-509):
 *  Lparms int32[5] = {Nz, Nf, NT, DEM key, DDM key}; Rparms f64[3] = {area cm^2, f0 Hz, log10 step};
 *  Parms f64 (15,Nz) column-major; T/DEM/DDM ignored (NT must be 0); RL f64 (7,Nf) column-major.
 * Returns 0 ok, 1 bad sizes, 2 DEM/DDM requested (unsupported).
 */
int oracle_get_mw(const int32_t *Lparms, const double *Rparms, const double *Parms,
                  const double *T_arr, const double *DEM_arr, const double *DDM_arr, double *RL)
{
    (void)T_arr; (void)DEM_arr; (void)DDM_arr;
    const int Nz = Lparms[0], Nf = Lparms[1];
    if (Nz < 0 || Nf <= 0) return 1;
    if (Lparms[2] > 0) return 2;
    const double area = Rparms[0], f0 = Rparms[1], step = Rparms[2];
    for (int f = 0; f < Nf; ++f) {
        const double nu = f0 * pow(10.0, step * f);
        double I6[6] = {0, 0, 0, 0, 0, 0};
        voxel_t prev; int have_prev = 0;
        for (int k = 0; k < Nz; ++k) {
            const double *P = Parms + (size_t)k * 15;
            voxel_t vx;
            vx.dz = P[0]; vx.T = P[1]; vx.ne = P[2]; vx.B = P[3];
            vx.th = P[4] * M_PI / 180.0;
            const int flag = (int)P[6];
            vx.gr_on = !(flag & 1); vx.ff_on = !(flag & 2);
            vx.smax = (int)P[7];
            vx.scale = (P[14] > 0.0) ? P[14] / area : 1.0;
            if (!(vx.dz > 0.0) || !(vx.T > 0.0) || !(vx.ne > 0.0) || !(vx.B >= 0.0) ||
                !isfinite(vx.dz) || !isfinite(vx.T) || !isfinite(vx.ne) || !isfinite(vx.B) || !isfinite(vx.th)) {
                have_prev = 0;                               /* empty / invalid voxel: transparent, */
                continue;                                    /* and no layer is placed across it    */
            }
            vx.cth = cos(vx.th); vx.sth = sin(vx.th);
            if (have_prev && vx.B > 0.0 && prev.B > 0.0) between(I6, nu, &prev, &vx);
            const int x_to_R = vx.cth >= 0.0;
            for (int sg = -1; sg <= 1; sg += 2) {
                mode_t m;
                mode_eval(nu, vx.ne, vx.B, vx.T, vx.cth, vx.sth, sg, vx.ff_on, &m, 0, 0);
                const int to_R = (sg < 0) ? x_to_R : !x_to_R;
                apply_mode(I6, to_R, m.prop, m.kap * vx.dz, m.src * vx.scale);
            }
            prev = vx; have_prev = 1;
        }
        double *o = RL + (size_t)f * 7;
        const double to_sfu = area / (AU_CM * AU_CM) / SFU;
        o[0] = nu / 1e9;
        for (int q = 0; q < 6; ++q) o[1 + q] = I6[q] * to_sfu;
    }
    return 0;
}

/*
 * Batched twin with the fastGRFF get_mw_slice layout (script/resample_with_ray_tracing.py:404-446):
 *  Lparms_M int32[6] = {Npix, Nz, Nf, NT, DEM key, DDM key}; Rparms_M f64 (3,Npix);
 *  Parms_M f64 (15,Nz,Npix); RL_M f64 (7,Nf,Npix); all Fortran order; status int32 (Npix).
 */
int oracle_get_mw_slice(const int32_t *Lparms_M, const double *Rparms_M, const double *Parms_M,
                        const double *T_arr, const double *DEM_arr, const double *DDM_arr,
                        double *RL_M, int32_t *status)
{
    const int Npix = Lparms_M[0], Nz = Lparms_M[1], Nf = Lparms_M[2];
    if (Npix < 0 || Nz < 0 || Nf <= 0) return 1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int p = 0; p < Npix; ++p) {
        /* NT slot of the batched header is 1 with dummy arrays at the reference call site */
        const int32_t L[5] = {Nz, Nf, 0, Lparms_M[4], Lparms_M[5]};
        status[p] = oracle_get_mw(L, Rparms_M + (size_t)p * 3, Parms_M + (size_t)p * 15 * Nz,
                                  T_arr, DEM_arr, DDM_arr, RL_M + (size_t)p * 7 * Nf);
    }
    return 0;
}
