/*
 * fep_beutler_kernel.cuh -- the Beutler soft-core path as ONE kernel per step: the pass at the current
 * lambda (forces, shift forces, Vc/Vv, dV/dlambda) and the energy-only foreign-lambda passes share
 * one load of every pair, one evaluation of everything that does not depend on lambda, and one
 * launch.  This is the path taken by all BASELINE.json configurations that use the Beutler
 * soft-core without a potential switch; Gapsys, no-soft-core and pot-switch go through the generic
 * kernels of fep_kernels.cu.
 *
 * What is computed (reference, src/gromacs):
 *   FORCE part    gmxlib/nonbonded/nb_free_energy.cpp:466-1179 with computeForces == true
 *   foreign part  nbnxm/freeenergydispatch.cpp:236-306 calling the energy-only flavour once per
 *                 lambda point
 * How (ours):
 *   - flat pair space, a warp owns 32 consecutive pair slots per trip, several trips per thread;
 *     one 16-byte pair record (fep_types.h) -> atom data -> type table is the whole load chain;
 *   - per state the interaction is expressed with coefficients that are ZERO when the state, the
 *     charge product, the LJ parameters or the lambda-independent part of a cut-off test rule the
 *     term out: straight-line FMA/MUFU code instead of the reference's masks and branches; a
 *     state nobody in the warp needs is skipped for 32 pairs at once;
 *   - foreign lambda: sums over pairs are kept per state because every lambda dependence outside
 *     the soft-core radius is a weight applied after the sum:
 *         E(p)         = sum_s lfacC[s][p] (C_s + Cp_s[p]) + lfacV[s][p] (G_s + V_s[p])
 *         dVdl_coul(p) = (C_B + Cp_B[p]) - (C_A + Cp_A[p]),  dVdl_vdw(p) = (G_B + V_B[p]) - (G_A + V_A[p])
 *     C_s: RF / Ewald / exclusion terms linear in qq[s] (:1023-1101) plus the whole Coulomb energy
 *     of state s when alphaCoul == 0; G_s: LJ-PME grid term (:1103-1136); V_s[p], Cp_s[p]: LJ and
 *     Coulomb energy with the soft-core radius of point p (:804-971).  State-A sums and per-pair
 *     B-minus-A DIFFERENCES are accumulated, so pairs with identical end states cancel exactly,
 *     as they do in the reference where the difference is formed per pair (:1005-1020).
 *     (energy-only passes have no soft-core term in dV/dlambda: it is built from force terms,
 *     which are zero when computeForces == false, :754-755,1005-1013.)
 *   - the lambda factors arrive as a __grid_constant__ kernel parameter: constant-bank operands
 *     of the FMAs, no loads in the loop over lambda points;
 *   - forces leave through the atom-sorted scatter of fep_types.h (no atomics).
 *
 * MODE 0: alphaCoul == 0 (GROMACS default sc-coul = no): rC == r, per point and state
 *         d = alphaVdwEff sigma6 sclfacV + r^6 ; 1/d by MUFU.RCP ; LJ from 1/d : 8 instructions.
 * MODE 1: alphaCoul == alphaVdw and lambdaCoul == lambdaVdw at every point: one radius; the
 *         Coulomb part needs d^(-1/6) = ex2(-lg2(d)/6).
 * MODE 2: separate Coulomb and LJ radii.
 */
#ifndef FEPB200_FEP_BEUTLER_KERNEL_CUH
#define FEPB200_FEP_BEUTLER_KERNEL_CUH
#include <cstdlib>
#include <cstring>

#include "fep_front.cuh"

#define FULL_MASK 0xffffffffu

struct BeutlerStep
{
    /* current lambda (nb_free_energy.cpp:420-449) */
    float cur_lfc[2], cur_lfv[2], cur_sclc[2], cur_sclv[2], cur_scdlc[2], cur_scdlv[2];
    /* chunk of foreign lambda points */
    float sclv[2][FEP_FB_MAXC]; /* soft-core lambda factor, vdw, per state */
    float sclc[2][FEP_FB_MAXC]; /* same for coulomb                         */
    float lfc[2][FEP_FB_MAXC];  /* {1-lambda_c, lambda_c}                   */
    float lfv[2][FEP_FB_MAXC];
    int   p0, np;               /* first point of the chunk, valid points   */
    int   want_shift;           /* also store the trips' forces sorted by shift vector */
    int   per_trip_energy;      /* more than one energy-group pair: Vc/Vv per trip instead of per CTA */
    int   n_tiles, n_parts;     /* CTAs of this launch; stride of the rows of cta_part it writes */
    int   always_check;         /* a lambda outside [0,1]: no fast path     */
};

/* sums N8*8 per-lane values over the warp; afterwards lane l < 8 holds, for group g, the value
 * with index 8*g + 4*(l&1) + 2*((l>>1)&1) + ((l>>2)&1) */
template<int N8>
__device__ __forceinline__ void warp_sum_groups(float (&v)[N8 * 8], float (&out)[N8], int lane)
{
#pragma unroll
    for (int g = 0; g < N8; g++)
    {
        float a[4], b[2], c;
        {
            const bool up = lane & 1;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const float send = up ? v[8 * g + i] : v[8 * g + i + 4];
                const float keep = up ? v[8 * g + i + 4] : v[8 * g + i];
                a[i]             = keep + __shfl_xor_sync(FULL_MASK, send, 1);
            }
        }
        {
            const bool up = lane & 2;
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                const float send = up ? a[i] : a[i + 2];
                const float keep = up ? a[i + 2] : a[i];
                b[i]             = keep + __shfl_xor_sync(FULL_MASK, send, 2);
            }
        }
        {
            const bool  up   = lane & 4;
            const float send = up ? b[0] : b[1];
            const float keep = up ? b[1] : b[0];
            c                = keep + __shfl_xor_sync(FULL_MASK, send, 4);
        }
        c += __shfl_xor_sync(FULL_MASK, c, 8);
        c += __shfl_xor_sync(FULL_MASK, c, 16);
        out[g] = c;
    }
}

/* lambda-independent data of one state of one pair; every coefficient is zero when the term it
 * multiplies does not apply */
struct StateConsts
{
    float c6_6, c12_12, shiftc, kv, kc, qe, qsh, qkrf;
};

/* One foreign lambda point of one state: LJ energy vv (and Coulomb energy vc when the Coulomb
 * radius is soft-cored).  CHECK = false is the fast path for warps in which no lane needs the
 * lambda-dependent cut-off tests or the r^-6 clamp (see the caller): 5 instructions in MODE 0
 * (FFMA, MUFU.RCP, FFMA, FFMA + the caller's FADD); CHECK = true adds clamp, compare and select. */
template<bool EWALD, int MODE, bool CHECK>
__device__ __forceinline__ void fb_point(const StateConsts& st, float r6, float sclv, float sclc, float thr_v,
                                         float rcoulomb6, float& vv, float& vc)
{
    const float dv  = fmaf(st.kv, sclv, r6);
    float       ri6 = fep_rcp(dv);
    if (CHECK)
    {
        ri6 = fminf(ri6, FEP_MAX_RINV6);
    }
    vv = fmaf(ri6, fmaf(st.c12_12, ri6, -st.c6_6), st.shiftc);
    if (CHECK)
    {
        vv = dv < thr_v ? vv : 0.0f;
    }
    if (MODE != 0)
    {
        const float dc  = (MODE == 1) ? dv : fmaf(st.kc, sclc, r6);
        const float lg  = fep_lg2(dc);
        const float ric = fep_ex2(lg * (-1.0f / 6.0f));
        if (EWALD)
        {
            vc = fmaf(st.qe, ric, st.qsh);
        }
        else
        {
            const float rc2 = fep_ex2(lg * (1.0f / 3.0f));
            vc              = fmaf(st.qe, ric, fmaf(st.qkrf, rc2, st.qsh));
            if (CHECK)
            {
                vc = dc < rcoulomb6 ? vc : 0.0f;
            }
        }
    }
}

/* the loop over the C lambda points of a chunk for the states the warp needs */
template<bool EWALD, int MODE, int C, bool CHECK, bool DO_A, bool DO_B>
__device__ __forceinline__ void fb_points(const StateConsts (&st)[2], const BeutlerStep& bs, float r6, float thr_v,
                                          float rcoulomb6, float* acc)
{
#pragma unroll
    for (int p = 0; p < C; p++)
    {
        float vvA = 0.0f, vvB = 0.0f, vcA = 0.0f, vcB = 0.0f;
        if (DO_A)
        {
            fb_point<EWALD, MODE, CHECK>(st[0], r6, bs.sclv[0][p], bs.sclc[0][p], thr_v, rcoulomb6, vvA, vcA);
        }
        if (DO_B)
        {
            fb_point<EWALD, MODE, CHECK>(st[1], r6, bs.sclv[1][p], bs.sclc[1][p], thr_v, rcoulomb6, vvB, vcB);
        }
        if (DO_A)
        {
            acc[p] += vvA;
        }
        acc[C + p] += DO_A ? (DO_B ? vvB - vvA : -vvA) : vvB;
        if (MODE != 0)
        {
            if (DO_A)
            {
                acc[2 * C + p] += vcA;
            }
            acc[3 * C + p] += DO_A ? (DO_B ? vcB - vcA : -vcA) : vcB;
        }
    }
}

/* One state at the current lambda WITH forces (:747-1020).  Adds to the scalar force (already
 * multiplied by r^(p-2)), the lambda-weighted energies and dV/dlambda incl. the soft-core term. */
template<bool EWALD, int MODE>
__device__ __forceinline__ void fb_force_state(const StateConsts& st, const BeutlerStep& bs, int s, float r2, float r6,
                                               float r4, float rinv, float thr_v, float rcoulomb6, float krf,
                                               float crf, float sh_ewald, float& fscal, float& vctot, float& vvtot,
                                               float& dc, float& dv)
{
    const float sign = s == 0 ? -1.0f : 1.0f;
    /* Lennard-Jones with the soft-core radius rV^6 = alpha sigma6 sclfac + r^6 */
    const float d_v  = fmaf(st.kv, bs.cur_sclv[s], r6);
    const float rp_v = fep_rcp(d_v);
    const float ri6  = fminf(rp_v, FEP_MAX_RINV6);
    const float t12  = st.c12_12 * ri6;
    float       vv   = fmaf(ri6, t12 - st.c6_6, st.shiftc);
    float       fv   = ri6 * fmaf(12.0f, t12, -6.0f * st.c6_6); /* V12 - V6 */
    const bool  on_v = d_v < thr_v;
    vv               = on_v ? vv : 0.0f;
    fv               = on_v ? fv : 0.0f;
    /* F rV^-6 r^4, the two factors combined first (stays in fp32 range for hard cores at r -> 0) */
    fscal = fmaf(bs.cur_lfv[s] * fv, rp_v * r4, fscal);
    vvtot = fmaf(bs.cur_lfv[s], vv, vvtot);
    dv    = fmaf(sign, vv, dv);
    dv    = fmaf((bs.cur_lfv[s] * bs.cur_scdlv[s] * st.kv) * fv, rp_v, dv); /* (:1010-1012), kv = alphaEff sigma6 */

    if (MODE != 0)
    {
        /* soft-cored Coulomb radius (MODE 0: rC == r, done once for both states by the caller) */
        float       vc, fc;
        const float d_c  = (MODE == 1) ? d_v : fmaf(st.kc, bs.cur_sclc[s], r6);
        const float rp_c = (MODE == 1) ? rp_v : fep_rcp(d_c);
        const float lg   = fep_lg2(d_c);
        const float ric  = fep_ex2(lg * (-1.0f / 6.0f));
        if (EWALD)
        {
            vc = fmaf(st.qe, ric, st.qsh);
            fc = st.qe * ric;
        }
        else
        {
            const float rc2 = fep_ex2(lg * (1.0f / 3.0f));
            const float k2  = st.qkrf * rc2;
            vc              = fmaf(st.qe, ric, k2 + st.qsh);
            fc              = fmaf(st.qe, ric, -2.0f * k2);
            const bool on_c = d_c < rcoulomb6;
            vc              = on_c ? vc : 0.0f;
            fc              = on_c ? fc : 0.0f;
        }
        dc    = fmaf((bs.cur_lfc[s] * bs.cur_scdlc[s] * st.kc) * fc, rp_c, dc); /* (:1007-1009) */
        fscal = fmaf(bs.cur_lfc[s] * fc, rp_c * r4, fscal);
        vctot = fmaf(bs.cur_lfc[s], vc, vctot);
        dc    = fmaf(sign, vc, dc);
    }
}

template<int MODE, int C, bool FORCE>
struct AccLayout
{
    static constexpr int NPER = (MODE == 0) ? 2 : 4; /* per-point: V_A DV (Cp_A DCp)                */
    static constexpr int NFOR = C > 0 ? NPER * C + 4 : 0; /* + C_A DC G_A DG                         */
    static constexpr int NACC = NFOR + (FORCE ? 4 : 0);  /* + dV/dlambda coul, vdw, Vc, Vv at current lambda */
    static constexpr int N8   = (NACC + 7) / 8;
    static constexpr int iCA = NPER * C, iDC = iCA + 1, iGA = iCA + 2, iDG = iCA + 3;
    static constexpr int iCUR = NFOR;
    /* register budget: 4 CTAs of 128 threads per SM up to ~56 accumulators, else 2 */
    static constexpr int MINB = (NACC + (FORCE ? 10 : 0) > 56) ? 2 : (NACC > 30 ? 3 : (NACC > 8 ? 4 : 8));
};

extern __shared__ __align__(128) unsigned char fep_dyn_smem[];

/* ELEC: 0 = reaction field / plain cut-off, 1 = Ewald real space, 2 = Ewald + LJ-PME grid correction */
template<int ELEC, int MODE, int C, bool FORCE, bool STAGED>
__global__ void __launch_bounds__(FEP_FB_CTA, (AccLayout<MODE, C, FORCE>::MINB))
        fep_beutler_kernel(const __grid_constant__ KernelArgs ka, const __grid_constant__ BeutlerStep bs)
{
    using L             = AccLayout<MODE, C, FORCE>;
    constexpr bool EWALD = ELEC != 0;
    constexpr bool LJPME = ELEC == 2;
    constexpr int  N8    = L::N8 > 0 ? L::N8 : 1;
    constexpr int  NW    = FEP_FB_CTA / 32;
    __shared__ float  s_red[C > 0 ? NW : 1][N8 * 8];
    __shared__ double s_sum[N8 * 8];
    __shared__ __align__(8) unsigned long long s_bars[NW * FEP_RING_DEPTH];

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    /* through a shuffle, so that the compiler knows the value to be the same in all lanes (uniform registers for
     * the ring's addresses and the bulk copies) */
    const int warp = __shfl_sync(FULL_MASK, tid >> 5, 0);
    fep_pdl_launch_dependents(); /* the next kernel of the step may fill SM space we leave free */

    /* this warp's runs of trips; its ring of trip blocks in shared memory is filled by bulk copies */
    const FepWalk         walk  = fep_walk(ka, gridDim.x * NW, FORCE ? ka.run_trips : 1);
    /* warp-major numbering: the CTAs' first warps take the first gridDim.x runs, their second warps the next ones, ...:
     * when there are fewer runs than warps, the busy warps are spread evenly over the CTAs (and with them over the SMs) */
    FepCursor             issue = fep_cursor(ka, walk, warp * gridDim.x + blockIdx.x); /* next trip to bring into the ring */
    FepCursor             ahead = issue;                                        /* next trip to fetch the gathers of */
    const FepRing<STAGED> ring  = fep_ring_open<STAGED>(ka, walk, issue, fep_dyn_smem, s_bars, warp);

    /* Layout: [0,C) V_A, [C,2C) DV, MODE>0: [2C,3C) Cp_A, [3C,4C) DCp, then C_A DC G_A DG,
     * then (FORCE) dV/dlambda_coul, dV/dlambda_vdw, Vc, Vv of the current-lambda pass. */
    float acc[N8 * 8];
#pragma unroll
    for (int i = 0; i < N8 * 8; i++)
    {
        acc[i] = 0.0f;
    }
    FepSegment seg;
    fep_segment_clear(seg);
    const float thr_v = LJPME ? __int_as_float(0x7f800000) : ka.rvdw6; /* LJ-PME tests r, below */

    /* while trip q is evaluated, the gathers of trip q + 1 are in flight and the blocks of the two trips after it
     * are on their way into the ring */
    const unsigned int* tb = nullptr;
    FepFetch            nx;
    bool                more = fep_cursor_valid(ahead, walk);
    if (more)
    {
        tb = fep_ring_block<STAGED>(ring, 0, ahead.t);
        nx = fep_fetch<STAGED, LJPME>(ka, tb, lane);
        fep_cursor_next(ahead, walk);
    }
    for (int q = 0; more; q++)
    {
        const unsigned int* tb_cur = tb;
        const FepFetch      cur    = nx;
        /* the warp has finished with trip q - 1: its ring slot takes the block of trip q + FEP_RING_DEPTH - 1 */
        __syncwarp();
        if (fep_cursor_valid(issue, walk))
        {
            fep_ring_issue<STAGED>(ring, (q + FEP_RING_DEPTH - 1) & (FEP_RING_DEPTH - 1), issue.t);
        }
        fep_cursor_next(issue, walk);
        more = fep_cursor_valid(ahead, walk);
        if (more)
        {
            tb = fep_ring_block<STAGED>(ring, q + 1, ahead.t);
            nx = fep_fetch<STAGED, LJPME>(ka, tb, lane);
            fep_cursor_next(ahead, walk);
        }
        const FepSlot p = fep_slot<STAGED>(ka, tb_cur, cur, lane);

        float nfx, nfy, nfz, vctot = 0.0f, vvtot = 0.0f; /* nf = MINUS the force on the owner = force on the partner */
        {
            const float4 ta = p.ta, tb4 = p.tb;
            const float  qq[2]  = { p.qq[0], p.qq[1] };
            const float  c6g[2] = { ta.w, tb4.w };
            const bool   hard   = (ta.y > 0.0f && tb4.y > 0.0f); /* :597-628 */
            const float  a_v    = hard ? 0.0f : ka.alpha_v;
            const float  a_c    = hard ? 0.0f : ka.alpha_c;

            /* a padding slot is "a pair far beyond every cut-off": all of its terms vanish below without masks */
            FepPair     pr;
            const float r2 = p.active ? fmaxf(p.r2, FEP_MIN_RSQ) : 1.0e6f;
            pr.r2   = r2;
            pr.rinv = fep_rsqrt(r2);
            pr.r    = r2 * pr.rinv;
            const float r4   = r2 * r2;
            const float r6   = r4 * r2;
            const bool  incl = p.active && p.within && !p.excluded;
            /* lambda-independent parts of the interaction tests (:805-812, :880-890) */
            const bool  m_e = incl && ((EWALD || MODE == 0) ? pr.r < ka.rcoulomb : true);
            const bool  m_v = incl && (LJPME ? pr.r < ka.rvdw : true);

            float fscal = 0.0f, dcur_c = 0.0f, dcur_v = 0.0f;

            /* lambda-independent correction terms, linear in qq[s] / c6grid[s]; zero beyond the cut-offs
             * unless the pair is an exclusion (:1023-1136) */
            {
                float xc, fcorr, xv, fvcorr;
                fep_corrections<EWALD, FORCE, LJPME ? 1 : 0>(ka, pr, p.excluded, p.self, xc, fcorr, xv, fvcorr);
                if (C > 0)
                {
                    const float cA = qq[0] * xc, cB = qq[1] * xc;
                    acc[L::iCA] += cA;
                    acc[L::iDC] += cB - cA;
                    if (LJPME)
                    {
                        const float gA = c6g[0] * xv, gB = c6g[1] * xv;
                        acc[L::iGA] += gA;
                        acc[L::iDG] += gB - gA;
                    }
                }
                if (FORCE)
                {
                    const float q_all = fmaf(bs.cur_lfc[0], qq[0], bs.cur_lfc[1] * qq[1]);
                    vctot             = q_all * xc;
                    dcur_c            = (qq[1] - qq[0]) * xc;
                    fscal             = q_all * fcorr;
                    if (LJPME)
                    {
                        const float g_all = fmaf(bs.cur_lfv[0], c6g[0], bs.cur_lfv[1] * c6g[1]);
                        vvtot             = g_all * xv;
                        dcur_v            = (c6g[1] - c6g[0]) * xv;
                        fscal             = fmaf(g_all, fvcorr, fscal);
                    }
                }
            }

            /* per state: coefficients that are ZERO when the state's term does not apply (the state takes part if
             * any of qq, c6, c12 is non-zero, :747-752 -- a zero coefficient says the same) */
            StateConsts st[2];
            {
                const float4 tt[2] = { ta, tb4 };
#pragma unroll
                for (int s = 0; s < 2; s++)
                {
                    st[s].qe     = m_e ? qq[s] : 0.0f;
                    st[s].c6_6   = m_v ? tt[s].x * (1.0f / 6.0f) : 0.0f;
                    st[s].c12_12 = m_v ? tt[s].y * (1.0f / 12.0f) : 0.0f;
                    st[s].shiftc = st[s].c12_12 * ka.rep_cpot - st[s].c6_6 * ka.disp_cpot;
                    if (LJPME)
                    {
                        st[s].shiftc = fmaf(m_v ? c6g[s] : 0.0f, ka.sh_lj_ewald * (1.0f / 6.0f), st[s].shiftc);
                    }
                    st[s].kv = a_v * tt[s].z;
                    if (MODE != 0)
                    {
                        st[s].kc   = a_c * tt[s].z;
                        st[s].qsh  = EWALD ? -st[s].qe * ka.sh_ewald : -st[s].qe * ka.crf;
                        st[s].qkrf = st[s].qe * ka.krf;
                    }
                }
            }
            bool vdw_on[2]  = { st[0].c6_6 != 0.0f || st[0].c12_12 != 0.0f, st[1].c6_6 != 0.0f || st[1].c12_12 != 0.0f };
            bool elec_on[2] = { st[0].qe != 0.0f, st[1].qe != 0.0f };
            if (MODE == 0)
            {
                /* Coulomb radius not soft-cored: rC == r, so the Coulomb energy of a state is qq[s] times a
                 * lambda-independent function of r and the two states need one evaluation (:804-874) */
                const float k2 = EWALD ? -ka.sh_ewald : fmaf(ka.krf, r2, -ka.crf);
                const float u  = pr.rinv + k2;
                const float dq = st[1].qe - st[0].qe;
                if (C > 0)
                {
                    acc[L::iCA] = fmaf(st[0].qe, u, acc[L::iCA]);
                    acc[L::iDC] = fmaf(dq, u, acc[L::iDC]);
                }
                if (FORCE)
                {
                    const float qeff = fmaf(bs.cur_lfc[0], st[0].qe, bs.cur_lfc[1] * st[1].qe);
                    const float g    = EWALD ? pr.rinv : fmaf(-2.0f * ka.krf, r2, pr.rinv);
                    vctot            = fmaf(qeff, u, vctot);
                    dcur_c           = fmaf(dq, u, dcur_c);
                    fscal            = fmaf(qeff * g, pr.rinv * pr.rinv, fscal); /* F r^-6 r^4 */
                }
            }
            /* a state nobody in the warp needs is skipped; the choice is made once per 32 pairs */
            const bool needA = __any_sync(FULL_MASK, vdw_on[0] || (MODE != 0 && elec_on[0]));
            const bool needB = __any_sync(FULL_MASK, vdw_on[1] || (MODE != 0 && elec_on[1]));

            if (FORCE)
            {
                if (needA)
                {
                    fb_force_state<EWALD, MODE>(st[0], bs, 0, r2, r6, r4, pr.rinv, thr_v, ka.rcoulomb6, ka.krf, ka.crf,
                                                ka.sh_ewald, fscal, vctot, vvtot, dcur_c, dcur_v);
                }
                if (needB)
                {
                    fb_force_state<EWALD, MODE>(st[1], bs, 1, r2, r6, r4, pr.rinv, thr_v, ka.rcoulomb6, ka.krf, ka.crf,
                                                ka.sh_ewald, fscal, vctot, vvtot, dcur_c, dcur_v);
                }
                const float nfs = -fscal;
                nfx             = nfs * p.dx;
                nfy             = nfs * p.dy;
                nfz             = nfs * p.dz;
                acc[L::iCUR]     += dcur_c;
                acc[L::iCUR + 1] += dcur_v;
            }

            if (C > 0)
            {
                /* Per lane and state the lambda-dependent tests fall in one of three classes, because
                 * the soft-core radius satisfies r^6 <= rV^6 <= r^6 + alphaEff sigma6 for every lambda:
                 * always outside (r^6 >= rc^6: coefficients zeroed here), always inside, or borderline.
                 * Only warps with a borderline lane, or a lane so close that r^-6 needs its clamp,
                 * take the loop with the per-point tests. */
                bool slow = (r6 < 1.0e-15f && incl) || bs.always_check != 0;
#pragma unroll
                for (int s = 0; s < 2; s++)
                {
                    if (vdw_on[s])
                    {
                        if (r6 >= thr_v)
                        {
                            st[s].c6_6 = st[s].c12_12 = st[s].shiftc = 0.0f;
                            vdw_on[s]                                = false;
                        }
                        else
                        {
                            slow = slow || (r6 + st[s].kv >= thr_v);
                        }
                    }
                    if (MODE != 0 && !EWALD && elec_on[s])
                    {
                        if (r6 >= ka.rcoulomb6)
                        {
                            st[s].qe = st[s].qsh = st[s].qkrf = 0.0f;
                            elec_on[s]                        = false;
                        }
                        else
                        {
                            slow = slow || (r6 + st[s].kc >= ka.rcoulomb6);
                        }
                    }
                }
                const bool pA   = __any_sync(FULL_MASK, vdw_on[0] || (MODE != 0 && elec_on[0]));
                const bool pB   = __any_sync(FULL_MASK, vdw_on[1] || (MODE != 0 && elec_on[1]));
                const bool chk  = __any_sync(FULL_MASK, slow);
                if (!chk)
                {
                    if (pA && pB)
                    {
                        fb_points<EWALD, MODE, C, false, true, true>(st, bs, r6, thr_v, ka.rcoulomb6, acc);
                    }
                    else if (pA)
                    {
                        fb_points<EWALD, MODE, C, false, true, false>(st, bs, r6, thr_v, ka.rcoulomb6, acc);
                    }
                    else if (pB)
                    {
                        fb_points<EWALD, MODE, C, false, false, true>(st, bs, r6, thr_v, ka.rcoulomb6, acc);
                    }
                }
                else
                {
                    if (pA && pB)
                    {
                        fb_points<EWALD, MODE, C, true, true, true>(st, bs, r6, thr_v, ka.rcoulomb6, acc);
                    }
                    else if (pA)
                    {
                        fb_points<EWALD, MODE, C, true, true, false>(st, bs, r6, thr_v, ka.rcoulomb6, acc);
                    }
                    else if (pB)
                    {
                        fb_points<EWALD, MODE, C, true, false, true>(st, bs, r6, thr_v, ka.rcoulomb6, acc);
                    }
                }
            }
        }

        if (FORCE)
        {
            if (p.active)
            {
                /* the partner receives -f: scattered to this pair's own slot in the atom-sorted
                 * buffer (unique destination, no atomics; skipped pairs write their zero) */
                ka.fsorted[fep_tw<STAGED>(tb_cur + FEP_TW_DST + lane)] = make_float4(nfx, nfy, nfz, 0.0f);
            }
            /* the owner receives the sum over its segment of trips: per-lane sums now, one warp reduction at
             * the segment's last trip */
            seg.fx += nfx;
            seg.fy += nfy;
            seg.fz += nfz;
            if (bs.per_trip_energy)
            {
                /* several energy-group pairs: Vc/Vv go to the segment's pair */
                seg.vc += vctot;
                seg.vv += vvtot;
            }
            else
            {
                acc[L::iCUR + 2] += vctot;
                acc[L::iCUR + 3] += vvtot;
            }
            if (fep_tw<STAGED>(tb_cur + FEP_TH_FLAGS) & FEP_TRIP_LAST)
            {
                fep_segment_flush<STAGED>(ka, tb_cur, cur.head, seg, bs.want_shift != 0, bs.per_trip_energy != 0, lane);
            }
        }
    }

    if (L::NACC == 0)
    {
        fep_pdl_wait();
        return;
    }
    if (C == 0)
    {
        /* force-only pass: four sums per WARP (dV/dlambda coul, vdw; Vc, Vv -- the latter two only meaningful
         * with one energy-group pair), no CTA-wide step: a warp that is done leaves */
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            const float w = fep_warp_sum(acc[L::iCUR + k]);
            v             = lane == k ? w : v;
        }
        if (lane < 4)
        {
            ka.cta_part[(size_t)lane * bs.n_parts + (blockIdx.x * NW + warp)] = (double)v;
        }
        fep_pdl_wait();
        return;
    }
    float red[N8];
    warp_sum_groups<N8>(acc, red, lane);
    if (lane < 8)
    {
        const int k = 4 * (lane & 1) + 2 * ((lane >> 1) & 1) + ((lane >> 2) & 1);
#pragma unroll
        for (int g = 0; g < N8; g++)
        {
            s_red[C > 0 ? warp : 0][8 * g + k] = red[g];
        }
    }
    __syncthreads();
    if (tid < L::NACC)
    {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < (C > 0 ? NW : 1); w++)
        {
            s += (double)s_red[w][tid];
        }
        s_sum[tid] = s;
    }
    __syncthreads();
    if (FORCE && tid < 4)
    {
        /* rows: dV/dlambda coul, vdw; Vc, Vv (the latter two only meaningful with one energy-group pair) */
        ka.cta_part[(size_t)tid * bs.n_parts + blockIdx.x] = s_sum[L::iCUR + tid];
    }
    if (C > 0 && tid < bs.np)
    {
        const int    p  = tid;
        const double CA = s_sum[L::iCA] + (MODE != 0 ? s_sum[2 * C + p] : 0.0);
        const double DC = s_sum[L::iDC] + (MODE != 0 ? s_sum[3 * C + p] : 0.0);
        const double GA = s_sum[L::iGA] + s_sum[p];
        const double DG = s_sum[L::iDG] + s_sum[C + p];
        /* E = lfacC[A] C_A + lfacC[B] C_B + lfacV[A] G_A + lfacV[B] G_B with X_B = X_A + DX */
        const double e = (double)bs.lfc[0][p] * CA + (double)bs.lfc[1][p] * (CA + DC) + (double)bs.lfv[0][p] * GA
                         + (double)bs.lfv[1][p] * (GA + DG);
        const size_t o = (size_t)(3 * (bs.p0 + p)) * bs.n_tiles + blockIdx.x;
        ka.for_part[o]                  = e;
        ka.for_part[o + bs.n_tiles]     = DC;
        ka.for_part[o + 2 * bs.n_tiles] = DG;
    }
    /* nothing here depends on the preceding kernel; completing after it keeps the chain ordered */
    fep_pdl_wait();
}

/* ------------------------------------------------------------------------------------------- */
/* occ != nullptr: only report how many CTAs of this instantiation fit on one SM */
/* FEPB200_STAGE=direct: the A/B variant that reads the tile's records from global memory in the loop
 * instead of staging them through shared memory with bulk copies (profiles/) */
static bool fb_staged()
{
    static const bool staged = [] {
        const char* e = std::getenv("FEPB200_STAGE");
        return !(e && std::strcmp(e, "direct") == 0);
    }();
    return staged;
}

template<int ELEC, int MODE, int C, bool FORCE>
static void launch_one(const KernelArgs& ka, const BeutlerStep& bs, cudaStream_t stream, int* occ, bool chained)
{
    if (occ)
    {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, fep_beutler_kernel<ELEC, MODE, C, FORCE, true>, FEP_FB_CTA,
                                                      fep_ring_bytes(FEP_FB_CTA / 32));
        return;
    }
    if (fb_staged())
    {
        fep_launch_kernel_smem(fep_beutler_kernel<ELEC, MODE, C, FORCE, true>, dim3(bs.n_tiles), dim3(FEP_FB_CTA),
                               fep_ring_bytes(FEP_FB_CTA / 32), stream, chained, ka, bs);
    }
    else
    {
        fep_launch_kernel_smem(fep_beutler_kernel<ELEC, MODE, C, FORCE, false>, dim3(bs.n_tiles), dim3(FEP_FB_CTA), 0,
                               stream, chained, ka, bs);
    }
}

template<int ELEC, int MODE, bool FORCE>
static bool launch_size(const KernelArgs& ka, const BeutlerStep& bs, int c, cudaStream_t stream, int* occ, bool chained)
{
    switch (c)
    {
#define FEP_FB_CASE(N) \
    case N: launch_one<ELEC, MODE, N, FORCE>(ka, bs, stream, occ, chained); return true;
        FEP_FB_CASE(1)
        FEP_FB_CASE(2)
        FEP_FB_CASE(3)
        FEP_FB_CASE(4)
        FEP_FB_CASE(6)
        FEP_FB_CASE(7)
        FEP_FB_CASE(8)
        FEP_FB_CASE(11)
        FEP_FB_CASE(14)
        FEP_FB_CASE(16)
        FEP_FB_CASE(21)
        FEP_FB_CASE(24)
#undef FEP_FB_CASE
        case 0:
            if (FORCE)
            {
                launch_one<ELEC, MODE, 0, true>(ka, bs, stream, occ, chained);
                return true;
            }
            return false;
        default: return false;
    }
}

/* One translation unit per (EWALD, MODE): fep_beutler_inst.cu compiled with -DFB_EWALD=.. -DFB_MODE=.. defines this
 * function for its pair; fep_beutler.cu dispatches.  c = points per chunk (0: force-only pass); returns false for a
 * chunk size that is not instantiated.  occ != nullptr: only report how many CTAs fit on one SM. */
#define FB_INST_NAME2(E, M) fb_launch_e##E##_m##M
#define FB_INST_NAME(E, M) FB_INST_NAME2(E, M)
#define FB_INST_DECL(E, M) \
    bool FB_INST_NAME(E, M)(const KernelArgs& ka, const BeutlerStep& bs, int c, bool force, cudaStream_t stream, int* occ, bool chained)
#endif
