/*
 * oracle/fep_oracle.c -- TEST INFRASTRUCTURE: a CPU restatement (plain C, double precision)
 * of the reference's perturbed-pair free-energy kernel.  It is the checker the CUDA path is
 * compared with; it is never linked into, imported by or shipped with the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it (and the latter only when oracle/_ref is absent).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the 72 golden vectors of the reference's own unit test
 *       (src/gromacs/gmxlib/nonbonded/tests/refdata/*.xml -> tests/golden/nb_free_energy_kat.json),
 *   (2) the reference kernel itself compiled in place (oracle/_ref/libfepref_dp.so) on seeded
 *       random problems that exercise everything the 72 cases do not (k_rf != 0, sc-power 2,
 *       lambda_coul != lambda_vdw, many entries / energy groups / shift vectors, clamps,
 *       force-only and energy-only passes, the foreign-lambda loop).
 *
 * What is restated (all paths relative to /root/reference/src/gromacs):
 *   gmxlib/nonbonded/nb_free_energy.cpp:274-1187   the pair mathematics
 *   gmxlib/nonbonded/nb_free_energy.cpp:1315-1448  the selection of the soft-core flavour and of
 *                                                  scLambdasOrAlphasDiffer
 *   gmxlib/nonbonded/nb_softcore.h:45-279          Gapsys force-linearised forms
 *   nbnxm/freeenergydispatch.cpp:147-308           the per-step driver: one pass at the current
 *                                                  lambda, then L+1 energy-only foreign passes
 *   nbnxm/pairlist.cpp:2786-2838                   the list split over threads
 * The reference evaluates erf-based Ewald corrections with rational minimax fits
 * (simd/.../scalar_math.h:385-469 and the double versions); here they are evaluated from their
 * closed forms, erf(z)/z and 2exp(-z^2)/(sqrt(pi) z^2) - erf(z)/z^3 (documented in
 * simd/include/gromacs/simd/simd_math.h:1560-1610), with a Taylor series for small z, which
 * agrees with the double-precision fits to ~1e-14.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* flag bits, gmxlib/nonbonded/nonbonded.h:38-42 */
#define DO_FORCE (1 << 1)
#define DO_SHIFTFORCE (1 << 2)
#define DO_FOREIGNLAMBDA (1 << 3)
#define DO_POTENTIAL (1 << 4)

#define NUM_SHIFT 45
#define LAMBDA_COUL 2
#define LAMBDA_VDW 3

/* md_enums.h */
enum { EEL_CUT = 0, EEL_RF = 1, EEL_GRF = 2, EEL_PME = 3, EEL_EWALD = 4, EEL_P3M = 5,
       EEL_RF_NEC = 11, EEL_PME_USER = 13, EEL_PME_SWITCH = 14, EEL_PME_USERSWITCH = 15, EEL_RFZERO = 16 };
enum { VDW_PME = 5 };
enum { MOD_POTSWITCH = 3 };
enum { SC_BEUTLER = 0, SC_GAPSYS = 1 };
enum { KSC_BEUTLER, KSC_GAPSYS, KSC_NONE };

/* same fields as fepb200_params (include/fepb200.h), reals in double */
typedef struct fep_oracle_params
{
    int    eeltype, vdwtype, vdw_modifier;
    double epsfac, rcoulomb, rvdw, rvdw_switch, krf, crf;
    double sh_ewald, sh_lj_ewald, ewaldcoeff_q, ewaldcoeff_lj, dispersion_cpot, repulsion_cpot;
    int    softcoreType;
    double alphaVdw, alphaCoulomb;
    int    lambdaPower;
    double sigma6WithInvalidSigma, sigma6Minimum, gapsysScaleVdW, gapsysScaleCoul, gapsysSigma6VdW;
} fep_oracle_params;

/* nb_free_energy.cpp:99,107 */
static const double MIN_RSQ    = 1.0e-12;
static const double MAX_RINV6  = 1.0e15;
static const double TWO_RSQRTPI = 1.1283791670955125739; /* 2/sqrt(pi) */

/* erf(z)/z as a function of z^2 (what pmePotentialCorrection approximates) */
static double pme_potential_correction(double z2)
{
    if (z2 < 1e-4)
    {
        /* 2/sqrt(pi) * sum (-1)^n z^2n / (n! (2n+1)) */
        return TWO_RSQRTPI * (1.0 - z2 / 3.0 + z2 * z2 / 10.0 - z2 * z2 * z2 / 42.0);
    }
    double z = sqrt(z2);
    return erf(z) / z;
}

/* 2exp(-z^2)/(sqrt(pi) z^2) - erf(z)/z^3 (what pmeForceCorrection approximates) */
static double pme_force_correction(double z2)
{
    if (z2 < 1.0)
    {
        /* 2/sqrt(pi) * sum_{n>=1} (-1)^n z^(2n-2) / n! * 2n/(2n+1); cancellation-free */
        double sum = 0.0, term = 1.0; /* term = z^(2n-2)/n! */
        for (int n = 1; n <= 40; n++)
        {
            term = (n == 1) ? 1.0 : term * z2 / n;
            double c = (2.0 * n) / (2.0 * n + 1.0);
            sum += ((n & 1) ? -1.0 : 1.0) * term * c;
            if (term < 1e-20)
            {
                break;
            }
        }
        return TWO_RSQRTPI * sum;
    }
    double z = sqrt(z2);
    return TWO_RSQRTPI * exp(-z2) / z2 - erf(z) / (z2 * z);
}

/* nb_free_energy.cpp:121-163; DBL_EPSILON plays GMX_REAL_EPS of the double build */
static void pme_lj_correction(double rinv, double rsq, double coeff_sq, double coeff6_div6, int computed,
                              int self, double* pot, double* force)
{
    const double sw     = pow(8.0 * 2.2204460492503131e-16, 1.0 / 6.0);
    const double rinv2  = rinv * rinv;
    const double rinv6  = rinv2 * rinv2 * rinv2;
    const double x      = coeff_sq * (computed ? rsq : 0.0);
    const double e      = exp(-x);
    const double poly   = 1.0 + x + 0.5 * x * x;
    const double full   = rinv6 * (1.0 - e * poly);
    const double approx = coeff6_div6 * (1.0 + x * (-0.75 + 0.3 * x));
    const double term   = (x < sw) ? approx : full;
    *force              = (term - e * coeff6_div6) * rinv2;
    *pot                = self ? 0.5 * coeff6_div6 : term;
}

typedef struct
{
    /* per-call constants derived from params + lambda (nb_free_energy.cpp:319-449) */
    const fep_oracle_params* p;
    int    ksc, differ, elec_ewald, vdw_ewald, pot_switch, rf_type;
    double lfac_c[2], lfac_v[2], dlfac[2];
    double sclfac_c[2], sclfac_v[2], scdl_c[2], scdl_v[2];
    double sw_v3, sw_v4, sw_v5, sw_f2, sw_f3, sw_f4;
    double rcut_max2, lj_coeff_sq, lj_coeff6_div6;
} kernel_consts;

static int eel_is_rf(int e)
{
    return e == EEL_RF || e == EEL_GRF || e == EEL_RF_NEC || e == EEL_RFZERO;
}
static int eel_is_ewald(int e)
{
    return e == EEL_PME || e == EEL_EWALD || e == EEL_P3M || e == EEL_PME_USER || e == EEL_PME_SWITCH
           || e == EEL_PME_USERSWITCH;
}

static void setup_consts(kernel_consts* k, const fep_oracle_params* p, double lam_c, double lam_v)
{
    k->p = p;
    /* nb_free_energy.cpp:1324-1363 */
    if (p->softcoreType == SC_BEUTLER)
    {
        k->ksc = (p->alphaCoulomb == 0 && p->alphaVdw == 0) ? KSC_NONE : KSC_BEUTLER;
    }
    else
    {
        k->ksc = (p->gapsysScaleCoul == 0 && p->gapsysScaleVdW == 0) ? KSC_NONE : KSC_GAPSYS;
    }
    /* :1405-1419 */
    k->differ = 1;
    if (p->alphaCoulomb == 0 && p->alphaVdw == 0)
    {
        k->differ = 0;
    }
    else if (lam_c == lam_v && p->alphaCoulomb == p->alphaVdw)
    {
        k->differ = 0;
    }
    k->elec_ewald = eel_is_ewald(p->eeltype);
    k->vdw_ewald  = (p->vdwtype == VDW_PME);
    k->pot_switch = (p->vdw_modifier == MOD_POTSWITCH);
    k->rf_type    = (p->eeltype == EEL_CUT || eel_is_rf(p->eeltype)); /* :377-386 */

    k->lfac_c[0] = 1.0 - lam_c;
    k->lfac_c[1] = lam_c;
    k->lfac_v[0] = 1.0 - lam_v;
    k->lfac_v[1] = lam_v;
    k->dlfac[0]  = -1.0;
    k->dlfac[1]  = 1.0;
    for (int s = 0; s < 2; s++)
    {
        const double pw = p->lambdaPower;
        const double oc = 1.0 - k->lfac_c[s], ov = 1.0 - k->lfac_v[s];
        k->sclfac_c[s]  = (p->lambdaPower == 2) ? oc * oc : oc;
        k->sclfac_v[s]  = (p->lambdaPower == 2) ? ov * ov : ov;
        k->scdl_c[s]    = k->dlfac[s] * pw / 6.0 * ((p->lambdaPower == 2) ? oc : 1.0);
        k->scdl_v[s]    = k->dlfac[s] * pw / 6.0 * ((p->lambdaPower == 2) ? ov : 1.0);
    }
    if (k->pot_switch)
    {
        const double d = p->rvdw - p->rvdw_switch; /* :361-370 */
        k->sw_v3       = -10.0 / (d * d * d);
        k->sw_v4       = 15.0 / (d * d * d * d);
        k->sw_v5       = -6.0 / (d * d * d * d * d);
        k->sw_f2       = -30.0 / (d * d * d);
        k->sw_f3       = 60.0 / (d * d * d * d);
        k->sw_f4       = -30.0 / (d * d * d * d * d);
    }
    else
    {
        k->sw_v3 = k->sw_v4 = k->sw_v5 = k->sw_f2 = k->sw_f3 = k->sw_f4 = 0;
    }
    const double rmax = p->rcoulomb > p->rvdw ? p->rcoulomb : p->rvdw;
    k->rcut_max2      = rmax * rmax;
    k->lj_coeff_sq    = p->ewaldcoeff_lj * p->ewaldcoeff_lj;
    k->lj_coeff6_div6 = k->lj_coeff_sq * k->lj_coeff_sq * k->lj_coeff_sq / 6.0;
}

/* Gapsys: quadratic Coulomb (nb_softcore.h:45-69 used by :73-195).  Returns 1 when the
 * hard-core values have been replaced. */
static int gapsys_coulomb(const kernel_consts* k, double qq, double r, int s, double scale_eff, int ewald,
                          double* force, double* pot, double* dvdl_acc)
{
    const fep_oracle_params* p    = k->p;
    const double             lfac = k->lfac_c[s];
    if (!(lfac < 1.0 && 0.0 < scale_eff && p->epsfac != 0.0))
    {
        return 0;
    }
    const double lrev = 1.0 - lfac;
    double       rq   = sqrt(cbrt(lrev)) * (1.0 + fabs(qq / p->epsfac)) * scale_eff;
    const int    within_cut = (rq <= p->rcoulomb);
    if (p->rcoulomb < rq)
    {
        rq = p->rcoulomb;
    }
    if (!(r < rq))
    {
        return 0;
    }
    const double rinvq = 1.0 / rq;
    const double cst   = qq * rinvq;
    const double lin   = cst * r * rinvq;
    const double quad  = lin * r * rinvq;
    double       fq    = -2.0 * quad + 3.0 * lin;
    double       vq    = quad - 3.0 * (lin - cst);
    const double dq    = k->dlfac[s] * 0.5 * (lfac * (1.0 / lrev)) * (quad - 2.0 * lin + cst);
    if (ewald)
    {
        vq -= qq * p->sh_ewald; /* :184 */
    }
    else
    {
        fq -= qq * 2.0 * p->krf * r * r; /* :121-122 */
        vq += qq * (p->krf * r * r - p->crf);
    }
    *force = fq;
    *pot   = vq;
    if (within_cut)
    {
        *dvdl_acc += dq;
    }
    return 1;
}

/* Gapsys: quadratic LJ (nb_softcore.h:199-279) */
static int gapsys_lj(const kernel_consts* k, double c6, double c12, double r, double rsq, int s, double sigma6,
                     double scale_eff, double* force, double* pot, double* dvdl_acc)
{
    const fep_oracle_params* p    = k->p;
    const double             lfac = k->lfac_v[s];
    if (!(lfac < 1.0 && 0.0 < scale_eff))
    {
        return 0;
    }
    const double lrev = 1.0 - lfac;
    const double rq   = sqrt(cbrt(26.0 / 7.0 * sigma6 * lrev)) * scale_eff;
    if (!(r < rq))
    {
        return 0;
    }
    const double c6s = c6 / 6.0, c12s = c12 / 12.0;
    const double ri    = 1.0 / rq;
    const double ri6   = (ri * ri * ri) * (ri * ri * ri);
    const double ri7   = ri6 * ri;
    const double ri8   = ri7 * ri;
    const double t14   = c12s * ri7 * ri7 * rsq;
    const double t13   = c12s * ri7 * ri6 * r;
    const double t12   = c12s * ri6 * ri6;
    const double t8    = ri8 * c6s * rsq;
    const double t7    = ri7 * c6s * r;
    const double t6    = ri6 * c6s;
    const double quad  = 156.0 * t14 - 42.0 * t8;
    const double lin   = 168.0 * t13 - 48.0 * t7;
    const double cst   = 91.0 * t12 - 28.0 * t6;
    *force             = -quad + lin;
    *pot               = 0.5 * quad - lin + cst + (c12s * p->repulsion_cpot - c6s * p->dispersion_cpot);
    *dvdl_acc += k->dlfac[s] * 28.0 * (lfac * (1.0 / lrev))
                 * ((6.5 * t14 - t8) - (13.0 * t13 - 2.0 * t7) + (6.5 * t12 - t6));
    return 1;
}

/* One kernel pass over a list (scalar flavour of nb_free_energy.cpp:466-1179).
 * f / fshift may be NULL when forces are not requested. */
static void kernel_pass(const fep_oracle_params* p, int ntype, const double* nbfp, const double* nbfp_grid,
                        const double* x, const double* qA, const double* qB, const int* typeA,
                        const int* typeB, const double* shiftvec, int e0, int e1, const int* iinr,
                        const int* gid, const int* shift, const int* jindex, const int* jjnr,
                        const int* excl, int flags, double lam_c, double lam_v, double* f, double* fshift,
                        double* Vc, double* Vv, double* dvdl /*[2]*/)
{
    kernel_consts k;
    setup_consts(&k, p, lam_c, lam_v);
    const int do_f     = (flags & DO_FORCE) != 0;
    const int do_shift = (flags & DO_SHIFTFORCE) != 0;
    const int do_pot   = (flags & DO_POTENTIAL) != 0;
    double    dvdl_c = 0, dvdl_v = 0;

    for (int n = e0; n < e1; n++)
    {
        const int    ii = iinr[n], is = shift[n];
        const double ix = shiftvec[3 * is] + x[3 * ii], iy = shiftvec[3 * is + 1] + x[3 * ii + 1],
                     iz  = shiftvec[3 * is + 2] + x[3 * ii + 2];
        const double iqA = p->epsfac * qA[ii], iqB = p->epsfac * qB[ii];
        const int    ntiA = ntype * typeA[ii], ntiB = ntype * typeB[ii];
        double       vctot = 0, vvtot = 0, fix = 0, fiy = 0, fiz = 0;
        int          any = 0;

        for (int kk = jindex[n]; kk < jindex[n + 1]; kk++)
        {
            const int    jnr      = jjnr[kk];
            const int    included = (excl == NULL) || (excl[kk] != 0);
            const int    excluded = !included;
            const double dx = ix - x[3 * jnr], dy = iy - x[3 * jnr + 1], dz = iz - x[3 * jnr + 2];
            double       rsq    = dx * dx + dy * dy + dz * dz;
            const int    within = rsq < k.rcut_max2;
            if (!(within || excluded)) /* :667 */
            {
                continue;
            }
            any            = 1;
            const int self = (ii == jnr);

            /* :539-650 pair parameters */
            const int    tj[2] = { ntiA + typeA[jnr], ntiB + typeB[jnr] };
            const double qq[2] = { iqA * qA[jnr], iqB * qB[jnr] };
            double       c6[2], c12[2], c6grid[2], sig6[2], gsig6[2];
            for (int s = 0; s < 2; s++)
            {
                c6[s]     = nbfp[2 * tj[s]];
                c12[s]    = nbfp[2 * tj[s] + 1];
                c6grid[s] = k.vdw_ewald ? nbfp_grid[2 * tj[s]] : 0.0;
                if (c6[s] > 0 && c12[s] > 0)
                {
                    sig6[s]  = 0.5 * c12[s] / c6[s];
                    gsig6[s] = sig6[s];
                    if (sig6[s] < p->sigma6Minimum)
                    {
                        sig6[s] = p->sigma6Minimum;
                    }
                }
                else
                {
                    sig6[s]  = p->sigma6WithInvalidSigma;
                    gsig6[s] = p->gapsysSigma6VdW;
                }
            }
            const int    hard    = (c12[0] > 0 && c12[1] > 0);
            const double alpha_v = hard ? 0.0 : p->alphaVdw, alpha_c = hard ? 0.0 : p->alphaCoulomb;
            const double gscale_v = hard ? 0.0 : p->gapsysScaleVdW, gscale_c = hard ? 0.0 : p->gapsysScaleCoul;

            /* :722-741 */
            if (rsq < MIN_RSQ)
            {
                rsq = MIN_RSQ;
            }
            const double rinv = 1.0 / sqrt(rsq);
            const double r    = rsq * rinv;
            double       rp, rpm2;
            if (k.ksc == KSC_BEUTLER)
            {
                rpm2 = rsq * rsq;
                rp   = rpm2 * rsq;
            }
            else
            {
                rpm2 = rinv * rinv;
                rp   = 1.0;
            }

            double fscal = 0;

            if (included && within)
            {
                double vc[2] = { 0, 0 }, vv[2] = { 0, 0 }, fc[2] = { 0, 0 }, fv[2] = { 0, 0 };
                for (int s = 0; s < 2; s++)
                {
                    if (!(qq[s] != 0 || c6[s] != 0 || c12[s] != 0))
                    {
                        continue;
                    }
                    double rpinv_c, rinv_c, r_c, rpinv_v, rinv_v, r_v;
                    if (k.ksc == KSC_BEUTLER)
                    {
                        rpinv_c = 1.0 / (alpha_c * k.sclfac_c[s] * sig6[s] + rp);
                        r_c     = sqrt(cbrt(1.0 / rpinv_c)); /* = rpinv_c^(-1/6) */
                        rinv_c  = 1.0 / r_c;
                        if (k.differ)
                        {
                            rpinv_v = 1.0 / (alpha_v * k.sclfac_v[s] * sig6[s] + rp);
                            r_v     = sqrt(cbrt(1.0 / rpinv_v));
                            rinv_v  = 1.0 / r_v;
                        }
                        else
                        {
                            rpinv_v = rpinv_c;
                            rinv_v  = rinv_c;
                            r_v     = r_c;
                        }
                    }
                    else
                    {
                        rpinv_c = rpinv_v = 1.0;
                        rinv_c = rinv_v = rinv;
                        r_c = r_v = r;
                    }
                    /* Coulomb, :804-874 */
                    if ((k.elec_ewald ? r : r_c) < p->rcoulomb && qq[s] != 0)
                    {
                        if (k.elec_ewald)
                        {
                            vc[s] = qq[s] * (rinv_c - p->sh_ewald);
                            fc[s] = qq[s] * rinv_c;
                        }
                        else
                        {
                            vc[s] = qq[s] * (rinv_c + p->krf * r_c * r_c - p->crf);
                            fc[s] = qq[s] * (rinv_c - 2.0 * p->krf * r_c * r_c);
                        }
                        if (k.ksc == KSC_GAPSYS)
                        {
                            gapsys_coulomb(&k, qq[s], r_c, s, gscale_c, k.elec_ewald, &fc[s], &vc[s], &dvdl_c);
                        }
                    }
                    /* Van der Waals, :880-971 */
                    if ((k.vdw_ewald ? r : r_v) < p->rvdw && (c6[s] != 0 || c12[s] != 0))
                    {
                        double rinv6 = (k.ksc == KSC_BEUTLER) ? rpinv_v
                                                              : (rinv_v * rinv_v) * (rinv_v * rinv_v)
                                                                        * (rinv_v * rinv_v);
                        if (rinv6 > MAX_RINV6)
                        {
                            rinv6 = MAX_RINV6;
                        }
                        const double v6 = c6[s] * rinv6, v12 = c12[s] * rinv6 * rinv6;
                        vv[s] = (v12 + c12[s] * p->repulsion_cpot) / 12.0 - (v6 + c6[s] * p->dispersion_cpot) / 6.0;
                        fv[s] = v12 - v6;
                        if (k.ksc == KSC_GAPSYS)
                        {
                            gapsys_lj(&k, c6[s], c12[s], r, rsq, s, gsig6[s], gscale_v, &fv[s], &vv[s], &dvdl_v);
                        }
                        if (k.vdw_ewald)
                        {
                            vv[s] += c6grid[s] * p->sh_lj_ewald / 6.0;
                        }
                        if (k.pot_switch)
                        {
                            double d = r_v - p->rvdw_switch;
                            if (!(0.0 < d))
                            {
                                d = 0.0;
                            }
                            const double d2  = d * d;
                            const double sw  = 1.0 + d2 * d * (k.sw_v3 + d * (k.sw_v4 + d * k.sw_v5));
                            const double dsw = d2 * (k.sw_f2 + d * (k.sw_f3 + d * k.sw_f4));
                            fv[s]            = fv[s] * sw - r_v * vv[s] * dsw;
                            vv[s] *= sw;
                        }
                    }
                    /* :980-981 (the reference only does this when forces are computed; with
                     * computeForces == false fc/fv stay zero, which matters for dvdl below) */
                    if (do_f)
                    {
                        fc[s] *= rpinv_c;
                        fv[s] *= rpinv_v;
                    }
                    else
                    {
                        fc[s] = 0;
                        fv[s] = 0;
                    }
                }
                /* assemble, :986-1020 */
                for (int s = 0; s < 2; s++)
                {
                    vctot += k.lfac_c[s] * vc[s];
                    vvtot += k.lfac_v[s] * vv[s];
                    fscal += (k.lfac_c[s] * fc[s] + k.lfac_v[s] * fv[s]) * rpm2;
                    dvdl_c += vc[s] * k.dlfac[s];
                    dvdl_v += vv[s] * k.dlfac[s];
                    if (k.ksc == KSC_BEUTLER)
                    {
                        dvdl_c += k.lfac_c[s] * alpha_c * k.scdl_c[s] * fc[s] * sig6[s];
                        dvdl_v += k.lfac_v[s] * alpha_v * k.scdl_v[s] * fv[s] * sig6[s];
                    }
                }
            }

            /* excluded pairs with reaction field / plain cut-off, :1023-1054 */
            if (k.rf_type && excluded)
            {
                const double ff = -2.0 * p->krf;
                double       vv = p->krf * rsq - p->crf;
                if (self)
                {
                    vv *= 0.5;
                }
                for (int s = 0; s < 2; s++)
                {
                    vctot += k.lfac_c[s] * qq[s] * vv;
                    fscal += k.lfac_c[s] * qq[s] * ff;
                    dvdl_c += k.dlfac[s] * qq[s] * vv;
                }
            }
            /* Ewald real-space correction, :1056-1101 */
            if (k.elec_ewald && (excluded || r < p->rcoulomb))
            {
                const double beta = p->ewaldcoeff_q;
                const double z2   = rsq * beta * beta;
                double       v_lr = beta * pme_potential_correction(z2);
                double       f_lr = -z2 * beta * pme_force_correction(z2) * rinv * rinv;
                if (self)
                {
                    v_lr *= 0.5;
                }
                for (int s = 0; s < 2; s++)
                {
                    vctot -= k.lfac_c[s] * qq[s] * v_lr;
                    if (do_f)
                    {
                        fscal -= k.lfac_c[s] * qq[s] * f_lr;
                    }
                    dvdl_c -= k.dlfac[s] * qq[s] * v_lr;
                }
            }
            /* LJ-PME grid correction, :1103-1136 */
            if (k.vdw_ewald && (excluded || r < p->rvdw))
            {
                double v_lr, f_lr;
                pme_lj_correction(rinv, rsq, k.lj_coeff_sq, k.lj_coeff6_div6, 1, self, &v_lr, &f_lr);
                v_lr /= 6.0;
                for (int s = 0; s < 2; s++)
                {
                    vvtot += k.lfac_v[s] * c6grid[s] * v_lr;
                    if (do_f)
                    {
                        fscal += k.lfac_v[s] * c6grid[s] * f_lr;
                    }
                    dvdl_v += k.dlfac[s] * c6grid[s] * v_lr;
                }
            }

            if (do_f && fscal != 0)
            {
                const double tx = fscal * dx, ty = fscal * dy, tz = fscal * dz;
                fix += tx;
                fiy += ty;
                fiz += tz;
                f[3 * jnr] -= tx;
                f[3 * jnr + 1] -= ty;
                f[3 * jnr + 2] -= tz;
            }
        }

        if (any) /* :1151-1169 */
        {
            if (do_f)
            {
                f[3 * ii] += fix;
                f[3 * ii + 1] += fiy;
                f[3 * ii + 2] += fiz;
                if (do_shift)
                {
                    fshift[3 * is] += fix;
                    fshift[3 * is + 1] += fiy;
                    fshift[3 * is + 2] += fiz;
                }
            }
            if (do_pot)
            {
                Vc[gid[n]] += vctot;
                Vv[gid[n]] += vvtot;
            }
        }
    }
    dvdl[0] += dvdl_c;
    dvdl[1] += dvdl_v;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Same argument list as fepref_dispatch() in oracle/ref_build/harness.cpp (use_simd ignored). */
int fep_oracle_dispatch(const fep_oracle_params* p, int use_simd, int nthreads, int ntype, const double* nbfp,
                        const double* nbfp_grid, int natoms, const double* x, const double* qA,
                        const double* qB, const int* typeA, const int* typeB, const double* shiftvec, int nri,
                        const int* iinr, const int* gid, const int* shift, const int* jindex, const int* jjnr,
                        const int* excl, int ngrp, int flags, const double* lambda, int nforeign,
                        const double* all_lambda_coul, const double* all_lambda_vdw, double* f, double* fshift,
                        double* Vc, double* Vv, double* dvdl, double* foreign_e, double* foreign_dvdl,
                        int repeats, double* seconds)
{
    (void)use_simd;
    if (nthreads < 1)
    {
        nthreads = 1;
    }
    if (repeats < 1)
    {
        repeats = 1;
    }
    /* split entries over threads, rule of balance_fep_lists (pairlist.cpp:2786-2838) */
    int* first = (int*)calloc(nthreads + 1, sizeof(int));
    {
        const long tot = nri > 0 ? jindex[nri] : 0, target = (tot + nthreads - 1) / nthreads;
        int        dest = 0;
        long       have = 0;
        for (int n = 0; n < nri; n++)
        {
            const long nrj = jindex[n + 1] - jindex[n];
            if (dest + 1 < nthreads && have > 0 && have + nrj - target > target - have)
            {
                dest++;
                first[dest] = n;
                have        = 0;
            }
            have += nrj;
        }
        for (int t = dest + 1; t <= nthreads; t++)
        {
            first[t] = nri;
        }
    }
    const size_t fsz = (size_t)3 * natoms;
    double*      tf  = (double*)calloc((size_t)nthreads * fsz, sizeof(double));
    double*      tfs = (double*)calloc((size_t)nthreads * 3 * NUM_SHIFT, sizeof(double));
    double*      tvc = (double*)calloc((size_t)nthreads * ngrp, sizeof(double));
    double*      tvv = (double*)calloc((size_t)nthreads * ngrp, sizeof(double));
    double*      tdv = (double*)calloc((size_t)nthreads * 2, sizeof(double));
    const double lam_c = lambda[LAMBDA_COUL], lam_v = lambda[LAMBDA_VDW];
    double       best_f = 1e30, best_e = 1e30;

    for (int rep = 0; rep < repeats; rep++)
    {
        const double t0 = now_s();
        memset(tf, 0, (size_t)nthreads * fsz * sizeof(double));
        memset(tfs, 0, (size_t)nthreads * 3 * NUM_SHIFT * sizeof(double));
        memset(tvc, 0, (size_t)nthreads * ngrp * sizeof(double));
        memset(tvv, 0, (size_t)nthreads * ngrp * sizeof(double));
        memset(tdv, 0, (size_t)nthreads * 2 * sizeof(double));
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int t = 0; t < nthreads; t++)
        {
            kernel_pass(p, ntype, nbfp, nbfp_grid, x, qA, qB, typeA, typeB, shiftvec, first[t], first[t + 1],
                        iinr, gid, shift, jindex, jjnr, excl, flags & ~DO_FOREIGNLAMBDA, lam_c, lam_v,
                        tf + t * fsz, tfs + (size_t)t * 3 * NUM_SHIFT, tvc + (size_t)t * ngrp,
                        tvv + (size_t)t * ngrp, tdv + 2 * t);
        }
        memset(f, 0, fsz * sizeof(double));
        memset(fshift, 0, 3 * NUM_SHIFT * sizeof(double));
        memset(Vc, 0, ngrp * sizeof(double));
        memset(Vv, 0, ngrp * sizeof(double));
        dvdl[0] = dvdl[1] = 0;
        for (int t = 0; t < nthreads; t++)
        {
            if (flags & DO_FORCE)
            {
                for (size_t i = 0; i < fsz; i++)
                {
                    f[i] += tf[t * fsz + i];
                }
            }
            for (int i = 0; i < 3 * NUM_SHIFT; i++)
            {
                fshift[i] += tfs[(size_t)t * 3 * NUM_SHIFT + i];
            }
            for (int g = 0; g < ngrp; g++)
            {
                Vc[g] += tvc[(size_t)t * ngrp + g];
                Vv[g] += tvv[(size_t)t * ngrp + g];
            }
            dvdl[0] += tdv[2 * t];
            dvdl[1] += tdv[2 * t + 1];
        }
        const double dt = now_s() - t0;
        if (dt < best_f)
        {
            best_f = dt;
        }
    }

    if ((flags & DO_FOREIGNLAMBDA) && foreign_e)
    {
        /* freeenergydispatch.cpp:236-306 */
        const int kflags = (flags & ~(DO_FORCE | DO_SHIFTFORCE)) | DO_FOREIGNLAMBDA | DO_POTENTIAL;
        for (int rep = 0; rep < repeats; rep++)
        {
            const double t0 = now_s();
            for (int i = 0; i <= nforeign; i++)
            {
                const double lc = (i == 0) ? lam_c : all_lambda_coul[i - 1];
                const double lv = (i == 0) ? lam_v : all_lambda_vdw[i - 1];
                memset(tvc, 0, (size_t)nthreads * ngrp * sizeof(double));
                memset(tvv, 0, (size_t)nthreads * ngrp * sizeof(double));
                memset(tdv, 0, (size_t)nthreads * 2 * sizeof(double));
#pragma omp parallel for schedule(static) num_threads(nthreads)
                for (int t = 0; t < nthreads; t++)
                {
                    kernel_pass(p, ntype, nbfp, nbfp_grid, x, qA, qB, typeA, typeB, shiftvec, first[t],
                                first[t + 1], iinr, gid, shift, jindex, jjnr, excl, kflags, lc, lv, NULL, NULL,
                                tvc + (size_t)t * ngrp, tvv + (size_t)t * ngrp, tdv + 2 * t);
                }
                double e = 0, dc = 0, dv = 0;
                for (int t = 0; t < nthreads; t++)
                {
                    for (int g = 0; g < ngrp; g++)
                    {
                        e += tvc[(size_t)t * ngrp + g] + tvv[(size_t)t * ngrp + g];
                    }
                    dc += tdv[2 * t];
                    dv += tdv[2 * t + 1];
                }
                foreign_e[i]            = e;
                foreign_dvdl[2 * i]     = dc;
                foreign_dvdl[2 * i + 1] = dv;
            }
            const double dt = now_s() - t0;
            if (dt < best_e)
            {
                best_e = dt;
            }
        }
    }
    else
    {
        best_e = 0;
    }
    if (seconds)
    {
        seconds[0] = best_f;
        seconds[1] = best_e;
    }
    free(first);
    free(tf);
    free(tfs);
    free(tvc);
    free(tvv);
    free(tdv);
    return 0;
}

/* ============================================================================================
 * Perturbed 1-4 pair interactions: restatement of
 *   listed_forces/pairs.cpp:170-515  free_energy_evaluate_single<softcoreType>
 *   listed_forces/pairs.cpp:516-835  do_pairs_general, perturbed (bFreeEnergy) branch, F_LJ14
 *   pbcutil/pbc.cpp:825-851          pbc_dx_aiuc, rectangular boxes
 * The reference evaluates plain Coulomb and r^-6 / r^-12 through cubic-spline tables
 * (fr->pairsTable, made by make_tables(..., GMX_MAKETABLES_14ONLY)); here they are evaluated
 * analytically.  The reference's own test of this function accepts 1e-7 (double) / 1e-5 (float)
 * because of the tables; tests/test_oracle_pairs14.py pins this restatement to the golden vectors
 * of that test (listed_forces/tests/refdata/14Interaction_ListedForcesPairsTest_Ifunc_{0,1,2}.xml).
 * ========================================================================================== */
int fep_oracle_pairs14(const fep_oracle_params* p, double fudgeQQ, int npairs, const int* iatoms, const double* c6A,
                       const double* c12A, const double* c6B, const double* c12B, const double* x, const double* qA,
                       const double* qB, const double* box_diag, int pbc_type /* 0 none, 1 xyz, 2 xy */,
                       const int* gid, double lam_c, double lam_v, double* f, double* fshift, double* Vc, double* Vv,
                       double* dvdl /*[2]*/)
{
    int ksc;
    if (p->softcoreType == SC_BEUTLER)
    {
        ksc = (p->alphaCoulomb == 0 && p->alphaVdw == 0) ? KSC_NONE : KSC_BEUTLER; /* :721-762 */
    }
    else
    {
        ksc = (p->gapsysScaleCoul == 0 && p->gapsysScaleVdW == 0) ? KSC_NONE : KSC_GAPSYS;
    }
    /* :575-601 */
    const double LFC[2] = { 1.0 - lam_c, lam_c }, LFV[2] = { 1.0 - lam_v, lam_v }, DLF[2] = { -1.0, 1.0 };
    double       lfac_coul[2], lfac_vdw[2], dlfac_coul[2], dlfac_vdw[2];
    for (int i = 0; i < 2; i++)
    {
        const int p2  = p->lambdaPower == 2;
        lfac_coul[i]  = p2 ? (1 - LFC[i]) * (1 - LFC[i]) : (1 - LFC[i]);
        dlfac_coul[i] = DLF[i] * p->lambdaPower / 6.0 * (p2 ? (1 - LFC[i]) : 1);
        lfac_vdw[i]   = p2 ? (1 - LFV[i]) * (1 - LFV[i]) : (1 - LFV[i]);
        dlfac_vdw[i]  = DLF[i] * p->lambdaPower / 6.0 * (p2 ? (1 - LFV[i]) : 1);
    }
    for (int n = 0; n < npairs; n++)
    {
        const int itype = iatoms[3 * n], ai = iatoms[3 * n + 1], aj = iatoms[3 * n + 2];
        double    dx[3];
        int       ishift[3] = { 0, 0, 0 };
        for (int d = 0; d < 3; d++)
        {
            dx[d] = x[3 * ai + d] - x[3 * aj + d];
            if (pbc_type == 1 || (pbc_type == 2 && d < 2))
            {
                const double hbox = 0.5 * box_diag[d];
                if (dx[d] > hbox)
                {
                    dx[d] -= box_diag[d];
                    ishift[d]--;
                }
                else if (dx[d] <= -hbox)
                {
                    dx[d] += box_diag[d];
                    ishift[d]++;
                }
            }
        }
        const int    is = 5 * (3 * (ishift[2] + 1) + (ishift[1] + 1)) + (ishift[0] + 2);
        const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2];
        const double qq[2]  = { qA[ai] * qA[aj] * p->epsfac * fudgeQQ, qB[ai] * qB[aj] * p->epsfac * fudgeQQ };
        const double c6[2]  = { 6.0 * c6A[itype], 6.0 * c6B[itype] };
        const double c12[2] = { 12.0 * c12A[itype], 12.0 * c12B[itype] };
        const double rpm2 = r2 * r2, rp = rpm2 * r2, r = sqrt(r2);
        double       sigma6[2], gsig6[2];
        for (int i = 0; i < 2; i++)
        {
            if (c6[i] > 0 && c12[i] > 0)
            {
                sigma6[i] = 0.5 * c12[i] / c6[i];
                gsig6[i]  = sigma6[i];
                if (sigma6[i] < p->sigma6Minimum)
                {
                    sigma6[i] = p->sigma6Minimum;
                }
            }
            else
            {
                sigma6[i] = p->sigma6WithInvalidSigma;
                gsig6[i]  = p->gapsysSigma6VdW;
            }
        }
        const int    hard = (c12[0] > 0 && c12[1] > 0);
        const double a_c = hard ? 0 : p->alphaCoulomb, a_v = hard ? 0 : p->alphaVdw;
        const double g_c = hard ? 0 : p->gapsysScaleCoul, g_v = hard ? 0 : p->gapsysScaleVdW;
        double       fe[2] = { 0, 0 }, fv[2] = { 0, 0 }, ve[2] = { 0, 0 }, vv[2] = { 0, 0 }, de[2] = { 0, 0 },
               dv[2] = { 0, 0 };
        for (int i = 0; i < 2; i++)
        {
            if (!(qq[i] != 0 || c6[i] != 0 || c12[i] != 0))
            {
                continue;
            }
            double rpinv, r_coul, r_vdw, rQ = 0, rLJ = 0, scaleDvdl = 1;
            if (ksc == KSC_BEUTLER)
            {
                rpinv  = 1.0 / (a_c * lfac_coul[i] * sigma6[i] + rp);
                r_coul = pow(rpinv, -1.0 / 6.0);
            }
            else
            {
                rpinv  = 1.0 / rp;
                r_coul = r;
            }
            if (ksc == KSC_GAPSYS)
            {
                if (p->epsfac != 0 && LFC[i] < 1)
                {
                    rQ = pow(1.0 - LFC[i], 1.0 / 6.0) * (1.0 + fabs(qq[i] / p->epsfac)) * g_c;
                }
                if (rQ > p->rcoulomb)
                {
                    rQ        = p->rcoulomb;
                    scaleDvdl = 0;
                }
            }
            if (ksc == KSC_GAPSYS && r < rQ)
            {
                const double ri = 1.0 / rQ, cst = qq[i] * ri, lin = cst * r * ri, quad = lin * r * ri;
                fe[i] = (-2 * quad + 3 * lin) * rpinv;
                ve[i] = quad - 3 * (lin - cst);
                de[i] += scaleDvdl * DLF[i] * 0.5 * (LFC[i] / (1 - LFC[i])) * (quad - 2 * lin + cst);
            }
            else
            {
                ve[i] = qq[i] / r_coul;
                fe[i] = qq[i] / r_coul * rpinv;
            }
            if (ksc == KSC_BEUTLER)
            {
                rpinv = 1.0 / (a_v * lfac_vdw[i] * sigma6[i] + rp);
                r_vdw = pow(rpinv, -1.0 / 6.0);
            }
            else
            {
                rpinv = 1.0 / rp;
                r_vdw = r;
            }
            if (ksc == KSC_GAPSYS && LFV[i] < 1)
            {
                rLJ = pow(26.0 / 7.0 * gsig6[i] * (1.0 - LFV[i]), 1.0 / 6.0) * g_v;
            }
            if (ksc == KSC_GAPSYS && r < rLJ)
            {
                const double c6s = c6[i] / 6.0, c12s = c12[i] / 12.0, ri = 1.0 / rLJ;
                double       ri6 = ri * ri * ri;
                ri6 *= ri6;
                const double ri7 = ri6 * ri, ri8 = ri7 * ri;
                const double t14 = c12s * ri7 * ri7 * r2, t13 = c12s * ri7 * ri6 * r, t12 = c12s * ri6 * ri6;
                const double t8 = ri8 * c6s * r2, t7 = ri7 * c6s * r, t6 = ri6 * c6s;
                const double quad = 156 * t14 - 42 * t8, lin = 168 * t13 - 48 * t7, cst = 91 * t12 - 28 * t6;
                fv[i] = (-quad + lin) * rpinv;
                vv[i] = 0.5 * quad - lin + cst;
                dv[i] += DLF[i] * 28 * (LFV[i] / (1.0 - LFV[i])) * ((6.5 * t14 - t8) - (13 * t13 - 2 * t7) + (6.5 * t12 - t6));
            }
            else
            {
                const double ri6 = 1.0 / (r_vdw * r_vdw * r_vdw * r_vdw * r_vdw * r_vdw);
                const double v6 = c6[i] * ri6, v12 = c12[i] * ri6 * ri6;
                vv[i] = v12 / 12.0 - v6 / 6.0;
                fv[i] = (v12 - v6) * rpinv;
            }
        }
        double velec = 0, vvdw = 0, fscal = 0, dc = 0, dvv = 0;
        for (int i = 0; i < 2; i++)
        {
            velec += LFC[i] * ve[i];
            vvdw += LFV[i] * vv[i];
            fscal += (LFC[i] * fe[i] + LFV[i] * fv[i]) * rpm2;
            if (ksc == KSC_GAPSYS)
            {
                dc += de[i];
                dvv += dv[i];
            }
            dc += ve[i] * DLF[i];
            dvv += vv[i] * DLF[i];
            if (ksc == KSC_BEUTLER)
            {
                dc += LFC[i] * a_c * dlfac_coul[i] * fe[i] * sigma6[i];
                dvv += LFV[i] * a_v * dlfac_vdw[i] * fv[i] * sigma6[i];
            }
        }
        dvdl[0] += dc;
        dvdl[1] += dvv;
        Vc[gid[n]] += velec;
        Vv[gid[n]] += vvdw;
        for (int d = 0; d < 3; d++)
        {
            const double t = fscal * dx[d];
            f[3 * ai + d] += t;
            f[3 * aj + d] -= t;
            if (fshift && is != 22)
            {
                fshift[3 * is + d] += t;
                fshift[3 * 22 + d] -= t;
            }
        }
    }
    return 0;
}
