/*
 * fep_pair_math.cuh -- per-pair mathematics of the perturbed-pair kernel in fp32 for sm_100a.
 *
 * What is computed is specified by the reference CPU kernel
 *   src/gromacs/gmxlib/nonbonded/nb_free_energy.cpp:539-1136 (pair parameters, soft-core radii,
 *   Coulomb / Lennard-Jones per state, assembly over states A/B, excluded-pair reaction field,
 *   Ewald and LJ-PME real-space corrections) and nb_softcore.h:45-279 (Gapsys);
 * how it is computed is ours: pair-per-thread SIMT code, MUFU lg2/ex2/rcp/rsq for the sixth
 * roots, erff + a Taylor branch for the Ewald correction instead of the reference's rational
 * fits, everything lambda-independent hoisted out of the foreign-lambda loop.
 */
#ifndef FEPB200_FEP_PAIR_MATH_CUH
#define FEPB200_FEP_PAIR_MATH_CUH

#include "fep_types.h"

#define FEP_MIN_RSQ 1.0e-12f  /* nb_free_energy.cpp:99  */
#define FEP_MAX_RINV6 1.0e15f /* nb_free_energy.cpp:107 */

/* bare MUFU operations (no denormal / range fix-up code): every argument here is a normal number */
__device__ __forceinline__ float fep_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fep_rsqrt(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fep_lg2(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fep_ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

/* coordinates of compact atom k (packed xyz) */
__device__ __forceinline__ float3 fep_load_pos(const float* __restrict__ pos3, int k)
{
    const float* p = pos3 + 3 * (size_t)k;
    return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

/* x^(-1/6) for x > 0 (normal range): two MUFU ops */
__device__ __forceinline__ float fep_inv_sixth_root(float x)
{
    return fep_ex2(fep_lg2(x) * (-1.0f / 6.0f));
}

/* x^(1/6) */
__device__ __forceinline__ float fep_sixth_root(float x)
{
    return fep_ex2(fep_lg2(x) * (1.0f / 6.0f));
}

/* Ewald real-space correction (reference :109-119, 1056-1101).  With w = z^2 = beta^2 r^2:
 *   v_lr = beta * V(w),    V(w) = erf(z)/z
 *   f_lr = -beta^3 * B(w), B(w) = 2 exp(-w)/(sqrt(pi) w) - erf(z)/z^3 = 2 dV/dw
 * (f_lr is the factor that multiplies the distance vector).  Both are smooth functions of w, so for
 * w <= 16 (beta r <= 4: every pair inside a cut-off chosen with ewald-rtol >= 2e-8, and the
 * excluded neighbours beside it) they are evaluated branch-free as rational functions P(w)/Q(w),
 * like the reference does (simd_math.h pmeForceCorrection / pmePotentialCorrection); the
 * coefficients are our own weighted least-squares/Lawson fits (tools/fit_ewald_rational.py):
 * V [6/4] 6.9e-8, B [6/5] 6.3e-8 maximum relative error in exact arithmetic, 4e-7 (8.6e-8 rms)
 * evaluated with fp32 FMAs.  Beyond w = 16 (only possible for excluded pairs far outside the
 * cut-off) the closed forms are used. */
template<bool FORCE>
__device__ __forceinline__ void fep_ewald_correction(float r2, float r, float rinv, float beta, float beta2,
                                                     float beta3, float* v_lr, float* f_lr)
{
    const float w = beta2 * r2;
    if (w <= 16.0f)
    {
        float pv = 1.914866538e-08f;
        pv       = fmaf(pv, w, -1.938895923e-06f);
        pv       = fmaf(pv, w, 1.461872746e-04f);
        pv       = fmaf(pv, w, 3.943561305e-03f);
        pv       = fmaf(pv, w, 5.137383335e-02f);
        pv       = fmaf(pv, w, 2.060183881e-01f);
        pv       = fmaf(pv, w, 1.128379099e+00f);
        float qv = 1.009842853e-03f;
        qv       = fmaf(qv, w, 1.486172240e-02f);
        qv       = fmaf(qv, w, 1.175093691e-01f);
        qv       = fmaf(qv, w, 5.159104191e-01f);
        qv       = fmaf(qv, w, 1.0f);
        *v_lr    = beta * pv * fep_rcp(qv);
        if (FORCE)
        {
            float pb = -1.081098709e-08f;
            pb       = fmaf(pb, w, 1.028332784e-06f);
            pb       = fmaf(pb, w, -5.036981975e-05f);
            pb       = fmaf(pb, w, 2.020152613e-04f);
            pb       = fmaf(pb, w, -1.807793702e-02f);
            pb       = fmaf(pb, w, 3.588346990e-02f);
            pb       = fmaf(pb, w, -7.522528250e-01f);
            float qb = 1.302164612e-04f;
            qb       = fmaf(qb, w, 2.096415634e-03f);
            qb       = fmaf(qb, w, 2.163110569e-02f);
            qb       = fmaf(qb, w, 1.411150645e-01f);
            qb       = fmaf(qb, w, 5.523007151e-01f);
            qb       = fmaf(qb, w, 1.0f);
            *f_lr    = -beta3 * pb * fep_rcp(qb);
        }
    }
    else
    {
        const float z  = beta * r;
        const float ez = erff(z);
        *v_lr          = ez * rinv;
        if (FORCE)
        {
            const float iw = fep_rcp(w);
            *f_lr = -beta3 * (1.1283791671f * fep_ex2(-1.4426950408889634f * w) * iw - ez * iw * fep_rcp(z));
        }
    }
}

/* LJ-PME grid correction (reference :121-163, 1103-1136).  term = r^-6 (1 - exp(-x)(1+x+x^2/2)),
 * x = beta_lj^2 r^2; for x < 1 the series sum_{n>=3} (-1)^(n+1) (n-1)(n-2)/(2 n!) x^n is used
 * (times beta_lj^6 / x^3 instead of r^-6), which is accurate where the closed form cancels. */
template<bool FORCE>
__device__ __forceinline__ void fep_ljpme_correction(float r2, float rinv, float csq, float c6div6, bool self,
                                                     float* pot, float* force)
{
    const float rinv2 = rinv * rinv;
    const float x     = csq * r2;
    const float e     = fep_ex2(-1.4426950408889634f * x);
    float       term;
    if (x < 1.0f)
    {
        /* r^-6 = beta_lj^6 / x^3, so term = beta_lj^6 * sum_{n>=3} a_n x^(n-3) with
         * a_n = (-1)^(n+1) (n-1)(n-2)/(2 n!); below p = 6 * that sum (p(0) = 1), n = 14..3 */
        float p = -5.368308940e-09f;
        p       = fmaf(p, x, 6.359381359e-08f);
        p       = fmaf(p, x, -6.889329806e-07f);
        p       = fmaf(p, x, 6.764069264e-06f);
        p       = fmaf(p, x, -5.952380952e-05f);
        p       = fmaf(p, x, 4.629629630e-04f);
        p       = fmaf(p, x, -3.125000000e-03f);
        p       = fmaf(p, x, 1.785714286e-02f);
        p       = fmaf(p, x, -8.333333333e-02f);
        p       = fmaf(p, x, 3.000000000e-01f);
        p       = fmaf(p, x, -7.500000000e-01f);
        p       = fmaf(p, x, 1.000000000e+00f);
        term    = c6div6 * p;
    }
    else
    {
        const float rinv6 = rinv2 * rinv2 * rinv2;
        term              = rinv6 * (1.0f - e * (1.0f + x + 0.5f * x * x));
    }
    if (FORCE)
    {
        *force = (term - e * c6div6) * rinv2;
    }
    *pot = self ? 0.5f * c6div6 : term;
}

/* lambda-independent description of one pair */
struct FepPair
{
    float r2;    /* clamped at FEP_MIN_RSQ */
    float r, rinv;
    float r6;    /* Beutler only */
    float rpm2;  /* r^(p-2): r^4 with the Beutler radius power 6, r^-2 otherwise (:722-741) */
    float qq[2], c6[2], c12[2], sig6[2], c6g[2];
    float a_c, a_v;      /* effective Beutler alphas or Gapsys scales for this pair */
    float gbase[2];      /* Gapsys: (26/7 sigma6)^(1/6) per state */
    bool  nonzero[2];
    bool  included_within; /* included pair inside the cut-off sphere */
};

/* Included & within-cut-off part for one lambda point: adds to vc/vv totals, fscal and dvdl.
 * Mirrors reference :747-1020. */
template<int SC, bool EWALD, bool FORCE>
__device__ __forceinline__ void fep_included_terms(const KernelArgs& ka, const LambdaPoint& lp, const FepPair& pr,
                                                   float& vctot, float& vvtot, float& fscal, float& dvdl_c,
                                                   float& dvdl_v)
{
#pragma unroll
    for (int s = 0; s < 2; s++)
    {
        if (!pr.nonzero[s])
        {
            continue;
        }
        const float dlfac = s == 0 ? -1.0f : 1.0f;
        float       vc = 0.0f, vv = 0.0f, fc = 0.0f, fv = 0.0f;
        float       rpinv_c = 1.0f, rinv_c = pr.rinv, rpinv_v = 1.0f, rinv_v = pr.rinv;
        float       d_c = 0.0f, d_v = 0.0f;
        if (SC == FEP_SC_BEUTLER)
        {
            d_c     = fmaf(pr.a_c * lp.sclfac_c[s], pr.sig6[s], pr.r6);
            rpinv_c = fep_rcp(d_c);
            rinv_c  = fep_inv_sixth_root(d_c);
            if (lp.differ)
            {
                d_v     = fmaf(pr.a_v * lp.sclfac_v[s], pr.sig6[s], pr.r6);
                rpinv_v = fep_rcp(d_v);
                rinv_v  = fep_inv_sixth_root(d_v);
            }
            else
            {
                d_v     = d_c;
                rpinv_v = rpinv_c;
                rinv_v  = rinv_c;
            }
        }
        /* ---- Coulomb (:804-874) ---- */
        bool elec;
        if (EWALD)
        {
            elec = pr.r < ka.rcoulomb;
        }
        else
        {
            /* rC < rc  <=>  rC^6 < rc^6 for the soft-cored radius */
            elec = (SC == FEP_SC_BEUTLER) ? (d_c < ka.rcoulomb6) : (pr.r < ka.rcoulomb);
        }
        elec = elec && pr.qq[s] != 0.0f;
        if (elec)
        {
            float r_c = pr.r;
            if (EWALD)
            {
                vc = pr.qq[s] * (rinv_c - ka.sh_ewald);
                fc = pr.qq[s] * rinv_c;
            }
            else
            {
                if (SC == FEP_SC_BEUTLER)
                {
                    r_c = fep_rcp(rinv_c);
                }
                const float krf_r2 = ka.krf * r_c * r_c;
                vc                 = pr.qq[s] * (rinv_c + krf_r2 - ka.crf);
                fc                 = pr.qq[s] * (rinv_c - 2.0f * krf_r2);
            }
            if (SC == FEP_SC_GAPSYS)
            {
                /* nb_softcore.h:73-195 */
                const float lfac = lp.lfac_c[s];
                if (lfac < 1.0f && pr.a_c > 0.0f && ka.gapsys_facel != 0.0f)
                {
                    /* facel and the cut-off of the linearisation point are separate constants: for
                     * 1-4 pairs they are not the (fudged) epsfac / cut-off of the interaction itself
                     * (listed_forces/pairs.cpp:318-338) */
                    float      rq         = lp.g6_c[s] * (1.0f + fabsf(pr.qq[s] / ka.gapsys_facel)) * pr.a_c;
                    const bool within_cut = rq <= ka.gapsys_rcoul;
                    rq                    = fminf(rq, ka.gapsys_rcoul);
                    if (pr.r < rq)
                    {
                        const float rinvq = fep_rcp(rq);
                        const float cst   = pr.qq[s] * rinvq;
                        const float lin   = cst * pr.r * rinvq;
                        const float quad  = lin * pr.r * rinvq;
                        fc                = -2.0f * quad + 3.0f * lin;
                        vc                = quad - 3.0f * (lin - cst);
                        if (EWALD)
                        {
                            vc -= pr.qq[s] * ka.sh_ewald;
                        }
                        else
                        {
                            const float krf_r2 = ka.krf * pr.r * pr.r;
                            fc -= pr.qq[s] * 2.0f * krf_r2;
                            vc += pr.qq[s] * (krf_r2 - ka.crf);
                        }
                        if (within_cut)
                        {
                            dvdl_c += dlfac * 0.5f * lp.gdl_c[s] * (quad - 2.0f * lin + cst);
                        }
                    }
                }
            }
        }
        /* ---- Van der Waals (:880-971) ---- */
        bool vdw;
        if (SC == FEP_SC_BEUTLER)
        {
            vdw = ka.vdw_ewald ? (pr.r < ka.rvdw) : (d_v < ka.rvdw6);
        }
        else
        {
            vdw = pr.r < ka.rvdw;
        }
        vdw = vdw && (pr.c6[s] != 0.0f || pr.c12[s] != 0.0f);
        if (vdw)
        {
            float rinv6;
            if (SC == FEP_SC_BEUTLER)
            {
                rinv6 = rpinv_v;
            }
            else
            {
                const float ri2 = rinv_v * rinv_v;
                rinv6           = ri2 * ri2 * ri2;
            }
            rinv6           = fminf(rinv6, FEP_MAX_RINV6);
            const float v6  = pr.c6[s] * rinv6;
            const float v12 = pr.c12[s] * rinv6 * rinv6;
            vv = (v12 + pr.c12[s] * ka.rep_cpot) * (1.0f / 12.0f) - (v6 + pr.c6[s] * ka.disp_cpot) * (1.0f / 6.0f);
            fv = v12 - v6;
            if (SC == FEP_SC_GAPSYS)
            {
                /* nb_softcore.h:199-279 */
                const float lfac = lp.lfac_v[s];
                if (lfac < 1.0f && pr.a_v > 0.0f)
                {
                    const float rq = pr.gbase[s] * lp.g6_v[s] * pr.a_v;
                    if (pr.r < rq)
                    {
                        const float c6s = pr.c6[s] * (1.0f / 6.0f), c12s = pr.c12[s] * (1.0f / 12.0f);
                        const float ri  = fep_rcp(rq);
                        const float ri3 = ri * ri * ri;
                        const float ri6 = ri3 * ri3;
                        const float ri7 = ri6 * ri;
                        const float ri8 = ri7 * ri;
                        const float t14 = c12s * ri7 * ri7 * pr.r2;
                        const float t13 = c12s * ri7 * ri6 * pr.r;
                        const float t12 = c12s * ri6 * ri6;
                        const float t8  = ri8 * c6s * pr.r2;
                        const float t7  = ri7 * c6s * pr.r;
                        const float t6  = ri6 * c6s;
                        const float quad = 156.0f * t14 - 42.0f * t8;
                        const float lin  = 168.0f * t13 - 48.0f * t7;
                        const float cst  = 91.0f * t12 - 28.0f * t6;
                        fv               = -quad + lin;
                        vv = 0.5f * quad - lin + cst + (c12s * ka.rep_cpot - c6s * ka.disp_cpot);
                        dvdl_v += dlfac * 28.0f * lp.gdl_v[s]
                                  * ((6.5f * t14 - t8) - (13.0f * t13 - 2.0f * t7) + (6.5f * t12 - t6));
                    }
                }
            }
            if (ka.vdw_ewald)
            {
                vv += pr.c6g[s] * ka.sh_lj_ewald * (1.0f / 6.0f);
            }
            if (ka.pot_switch)
            {
                const float r_v  = (SC == FEP_SC_BEUTLER) ? fep_rcp(rinv_v) : pr.r;
                const float d    = fmaxf(r_v - ka.rvdw_switch, 0.0f);
                const float d2   = d * d;
                const float sw   = 1.0f + d2 * d * (ka.sw_v3 + d * (ka.sw_v4 + d * ka.sw_v5));
                if (FORCE)
                {
                    const float dsw = d2 * (ka.sw_f2 + d * (ka.sw_f3 + d * ka.sw_f4));
                    fv              = fv * sw - r_v * vv * dsw;
                }
                vv *= sw;
            }
        }
        /* ---- assemble (:980-1020) ---- */
        vctot = fmaf(lp.lfac_c[s], vc, vctot);
        vvtot = fmaf(lp.lfac_v[s], vv, vvtot);
        dvdl_c = fmaf(vc, dlfac, dvdl_c);
        dvdl_v = fmaf(vv, dlfac, dvdl_v);
        if (FORCE)
        {
            /* The reference scales by rC^-6 (:980-981) and later by r^(p-2) (:998-1002).  The two
             * factors are combined first: for hard-core pairs at r -> 0 the product F * rC^-6
             * leaves the fp32 range although the force itself does not. */
            fscal += lp.lfac_c[s] * (fc * (rpinv_c * pr.rpm2)) + lp.lfac_v[s] * (fv * (rpinv_v * pr.rpm2));
            if (SC == FEP_SC_BEUTLER && pr.a_v + pr.a_c != 0.0f)
            {
                /* the reference adds these only when forces are computed (:1005-1013 use the
                 * force terms, which stay zero in energy-only passes); they vanish for pairs that
                 * are not soft-cored (alpha_eff == 0) */
                dvdl_c += lp.lfac_c[s] * pr.a_c * lp.scdl_c[s] * (fc * rpinv_c) * pr.sig6[s];
                dvdl_v += lp.lfac_v[s] * pr.a_v * lp.scdl_v[s] * (fv * rpinv_v) * pr.sig6[s];
            }
        }
    }
}

/* Lambda-independent correction factors of one pair (reference :1023-1136):
 *   xc / fc multiply qq[s]  (excluded-pair reaction field, Ewald real-space correction)
 *   xv / fv multiply c6grid[s] (LJ-PME grid correction)                                   */
/* LJPME: 1 / 0 = LJ-PME known at compile time, -1 = ask ka.vdw_ewald */
template<bool EWALD, bool FORCE, int LJPME = -1>
__device__ __forceinline__ void fep_corrections(const KernelArgs& ka, const FepPair& pr, bool excluded, bool self,
                                                float& xc, float& fc, float& xv, float& fv)
{
    xc = fc = xv = fv = 0.0f;
    if (!EWALD)
    {
        if (ka.rf_type && excluded)
        {
            float vv = fmaf(ka.krf, pr.r2, -ka.crf);
            if (self)
            {
                vv *= 0.5f;
            }
            xc = vv;
            fc = -2.0f * ka.krf;
        }
    }
    else if (excluded || pr.r < ka.rcoulomb)
    {
        float v_lr, f_lr = 0.0f;
        fep_ewald_correction<FORCE>(pr.r2, pr.r, pr.rinv, ka.beta, ka.beta2, ka.beta3, &v_lr, &f_lr);
        if (self)
        {
            v_lr *= 0.5f;
        }
        xc = -v_lr;
        fc = -f_lr;
    }
    if ((LJPME < 0 ? ka.vdw_ewald != 0 : LJPME != 0) && (excluded || pr.r < ka.rvdw))
    {
        float pot, force = 0.0f;
        fep_ljpme_correction<FORCE>(pr.r2, pr.rinv, ka.lj_coeff_sq, ka.lj_coeff6_div6, self, &pot, &force);
        xv = pot * (1.0f / 6.0f);
        fv = force;
    }
}

#endif
