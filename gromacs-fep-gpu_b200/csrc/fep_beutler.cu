/*
 * fep_beutler.cu -- the Beutler soft-core path as ONE kernel per step: the pass at the current
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
#include "fep_beutler_kernel.cuh"

FB_INST_DECL(0, 0);
FB_INST_DECL(0, 1);
FB_INST_DECL(0, 2);
FB_INST_DECL(1, 0);
FB_INST_DECL(1, 1);
FB_INST_DECL(1, 2);
FB_INST_DECL(2, 0);
FB_INST_DECL(2, 1);
FB_INST_DECL(2, 2);

static bool launch_any(const KernelArgs& ka, const BeutlerStep& bs, bool ewald, int mode, int c, bool force, cudaStream_t stream,
                       int* occ, bool chained = false)
{
    /* LJ-PME implies PME electrostatics (the reference's grompp insists) */
    switch ((ewald ? (ka.vdw_ewald ? 6 : 3) : 0) + mode)
    {
        case 0: return fb_launch_e0_m0(ka, bs, c, force, stream, occ, chained);
        case 1: return fb_launch_e0_m1(ka, bs, c, force, stream, occ, chained);
        case 2: return fb_launch_e0_m2(ka, bs, c, force, stream, occ, chained);
        case 3: return fb_launch_e1_m0(ka, bs, c, force, stream, occ, chained);
        case 4: return fb_launch_e1_m1(ka, bs, c, force, stream, occ, chained);
        case 5: return fb_launch_e1_m2(ka, bs, c, force, stream, occ, chained);
        case 6: return fb_launch_e2_m0(ka, bs, c, force, stream, occ, chained);
        case 7: return fb_launch_e2_m1(ka, bs, c, force, stream, occ, chained);
        case 8: return fb_launch_e2_m2(ka, bs, c, force, stream, occ, chained);
        default: return false;
    }
}

static const int c_sizes[] = { 1, 2, 3, 4, 6, 7, 8, 11, 14, 16, 21, 24 };

extern "C" int fep_beutler_chunk_size(int n_points, int n_chunks_wanted)
{
    if (n_chunks_wanted < 1)
    {
        n_chunks_wanted = 1;
    }
    const int need = (n_points + n_chunks_wanted - 1) / n_chunks_wanted;
    for (int c : c_sizes)
    {
        if (c >= need)
        {
            return c;
        }
    }
    return FEP_FB_MAXC;
}

/* resident CTAs per SM of the kernel instantiation a launch would use (for one-wave tile sizing) */
extern "C" int fep_beutler_ctas_per_sm(int elec_ewald, int mode, int c, int force)
{
    KernelArgs  ka{};
    BeutlerStep bs{};
    int         occ = 0;
    const bool  ok = launch_any(ka, bs, elec_ewald != 0, mode, c, force != 0, nullptr, &occ);
    return (ok && occ > 0) ? occ : 1;
}

/* One step of the Beutler path: `do_force` -> the pass at the current lambda is computed (fused
 * with the first chunk of foreign points when `do_foreign`); `pts` are the HOST copies of the
 * lambda points, pts[0] = current.  Returns cudaSuccess (0) or an error code. */
extern "C" int fep_launch_beutler(const KernelArgs* kap, int elec_ewald, int mode, const LambdaPoint* cur,
                                  const LambdaPoint* pts, int do_force, int do_foreign, int want_shift,
                                  cudaStream_t stream, long long* counter, int chained)
{
    const KernelArgs& ka = *kap;
    const int         c  = do_foreign ? ka.chunk_points : 0;
    const int         np = do_foreign ? ka.n_points : 0;
    BeutlerStep       bs;
    for (int s = 0; s < 2; s++)
    {
        bs.cur_lfc[s]   = cur->lfac_c[s];
        bs.cur_lfv[s]   = cur->lfac_v[s];
        bs.cur_sclc[s]  = cur->sclfac_c[s];
        bs.cur_sclv[s]  = cur->sclfac_v[s];
        bs.cur_scdlc[s] = cur->scdl_c[s];
        bs.cur_scdlv[s] = cur->scdl_v[s];
    }
    bs.want_shift   = want_shift;
    bs.per_trip_energy = ka.n_gid > 1 ? 1 : 0;
    /* the fast path of the point loop assumes 0 <= soft-core lambda factor <= 1 */
    bs.always_check = 0;
    for (int q = 0; q < np; q++)
    {
        for (int s = 0; s < 2; s++)
        {
            if (!(pts[q].sclfac_v[s] >= 0.0f && pts[q].sclfac_v[s] <= 1.0f && pts[q].sclfac_c[s] >= 0.0f
                  && pts[q].sclfac_c[s] <= 1.0f))
            {
                bs.always_check = 1;
            }
        }
    }
    bs.n_tiles = do_foreign ? ka.n_tiles : ka.pass_n_tiles;
    /* the force-only pass writes its four sums per warp, the fused pass per CTA */
    bs.n_parts = do_foreign ? ka.n_tiles : ka.pass_n_tiles * (FEP_FB_CTA / 32);
    bool first    = true;
    for (int p0 = 0; p0 < np || first; p0 += (c > 0 ? c : 1))
    {
        bs.p0 = p0;
        bs.np = c > 0 ? ((np - p0 < c) ? np - p0 : c) : 0;
        for (int p = 0; p < FEP_FB_MAXC; p++)
        {
            /* padding points repeat the last valid one; their results are not written */
            const int q = c > 0 ? p0 + (p < bs.np ? p : bs.np - 1) : 0;
            for (int s = 0; s < 2; s++)
            {
                bs.sclv[s][p] = c > 0 ? pts[q].sclfac_v[s] : 0.0f;
                bs.sclc[s][p] = c > 0 ? pts[q].sclfac_c[s] : 0.0f;
                bs.lfc[s][p]  = c > 0 ? pts[q].lfac_c[s] : 0.0f;
                bs.lfv[s][p]  = c > 0 ? pts[q].lfac_v[s] : 0.0f;
            }
        }
        const bool force = first && do_force;
        const bool ok    = launch_any(ka, bs, elec_ewald != 0, mode, c, force, stream, nullptr, chained != 0);
        if (!ok)
        {
            return (int)cudaErrorInvalidValue;
        }
        (*counter)++;
        first   = false;
        chained = chained != 0 ? chained : (ka.pdl_chain ? 1 : 0); /* later chunks follow a kernel of this step */
        if (c == 0)
        {
            break;
        }
    }
    return (int)cudaGetLastError();
}
