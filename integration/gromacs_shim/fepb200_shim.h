/*
 * fepb200_shim.h -- the reference-side binding of libfepb200.so.
 *
 * This header is meant to be #included by the reference's
 *   src/gromacs/nbnxm/freeenergydispatch.cpp
 * (after its own includes and after the file-local haveSoftCore()); together with the six-line
 * hook of freeenergydispatch_fepb200.patch it routes FreeEnergyDispatch::dispatchFreeEnergyKernels
 * through the B200 library when the environment variable GMX_FEPB200 is set, leaving every other
 * line of GROMACS -- pair search, forcerec, interaction_const_t, mdrun -- as it is.
 *
 * The library is loaded with dlopen() (GMX_FEPB200_LIB or "libfepb200.so"), so the GROMACS build
 * system does not change.  Everything the CPU path passes to gmx_nb_free_energy_kernel is handed
 * over unchanged: the per-thread t_nblist objects of all localities (concatenated), coordinates,
 * shift vectors, nbfp tables, A/B charges and types, lambda, the foreign-lambda table.  Results are
 * added where FreeEnergyDispatch adds them (freeenergydispatch.cpp:236-306, :395-410).
 *
 * Mixed-precision build only (real == float).
 */
#ifndef FEPB200_SHIM_H
#define FEPB200_SHIM_H

#include "fepb200_shim_common.h"

namespace fepb200shim
{

/* What FreeEnergyDispatch::dispatchFreeEnergyKernels does, through libfepb200. */
inline void dispatch(const PairlistSets&                              pairlistSets,
                     const gmx::ArrayRefWithPadding<const gmx::RVec>& coords,
                     gmx::ForceWithShiftForces*                       forceWithShiftForces,
                     const int                                        ntype,
                     const interaction_const_t&                       ic,
                     gmx::ArrayRef<const gmx::RVec>                   shiftvec,
                     gmx::ArrayRef<const real>                        nbfp,
                     gmx::ArrayRef<const real>                        nbfp_grid,
                     gmx::ArrayRef<const real>                        chargeA,
                     gmx::ArrayRef<const real>                        chargeB,
                     gmx::ArrayRef<const int>                         typeA,
                     gmx::ArrayRef<const int>                         typeB,
                     t_lambda*                                        fepvals,
                     gmx::ArrayRef<const real>                        lambda,
                     gmx_enerdata_t*                                  enerd,
                     const gmx::StepWorkload&                         stepWork,
                     const bool                                       softCore)
{
    static_assert(sizeof(real) == sizeof(float), "the fepb200 shim needs a mixed-precision build");
    load(&ic);
    Api& a = api();

    const double t0 = now();
    /* Cadence of the reference hooks these calls replace: constants when they change
     * (PME tuning may move rcoulomb / ewaldcoeff_q), atoms and pair list on search steps
     * (constructPairlist, pairlist.cpp:4437-4468; atom properties change with the DD partitioning,
     * which happens on search steps only), lambdas when they change, coordinates every step. */
    const bool           search = stepWork.doNeighborSearch || a.calls == 0;
    const fepb200_params p      = toParams(ic);
    if (search || std::memcmp(&p, &a.lastParams, sizeof(p)) != 0)
    {
        check(a.set_params(a.ctx, &p), "set_params");
        a.lastParams = p;
    }
    if (search)
    {
        check(a.set_nbfp(a.ctx, ntype, nbfp.data(), nbfp_grid.empty() ? nullptr : nbfp_grid.data()), "set_nbfp");
        check(a.set_atoms(a.ctx, static_cast<int>(chargeA.size()), chargeA.data(), chargeB.data(), typeA.data(),
                          typeB.data()),
              "set_atoms");
    }
    const int          nLambda = fepvals->n_lambda;
    std::vector<float> allCoul(nLambda), allVdw(nLambda), lam(lambda.begin(), lambda.end());
    for (int i = 0; i < nLambda; i++)
    {
        allCoul[i] = static_cast<float>(fepvals->all_lambda[FreeEnergyPerturbationCouplingType::Coul][i]);
        allVdw[i]  = static_cast<float>(fepvals->all_lambda[FreeEnergyPerturbationCouplingType::Vdw][i]);
    }
    if (search || lam != a.lastLambda || allCoul != a.lastAllCoul || allVdw != a.lastAllVdw)
    {
        check(a.set_lambdas(a.ctx, lambda.data(), nLambda, allCoul.data(), allVdw.data()), "set_lambdas");
        a.lastLambda  = lam;
        a.lastAllCoul = allCoul;
        a.lastAllVdw  = allVdw;
    }

    const int numGroupPairs = enerd->grpp.nener;
    if (search)
    {
        /* the FEP pair lists of all localities and threads as they are: the library concatenates them on the device
         * (what the fork's combine_fep_lists does element by element on the host for its GPU path, pairlist.cpp:2867) */
        std::vector<fepb200_list_view> views;
        const int                      numLocalities = (pairlistSets.params().haveMultipleDomains ? 2 : 1);
        for (int l = 0; l < numLocalities; l++)
        {
            const auto lists = pairlistSets.pairlistSet(static_cast<gmx::InteractionLocality>(l)).fepLists();
            for (const auto& nl : lists)
            {
                if (nl->nri == 0)
                {
                    continue;
                }
                fepb200_list_view v;
                v.nri      = nl->nri;
                v.iinr     = nl->iinr.data();
                v.gid      = nl->gid.data();
                v.shift    = nl->shift.data();
                v.jindex   = nl->jindex.data();
                v.jjnr     = nl->jjnr.data();
                v.excl_fep = nl->excl_fep.data();
                views.push_back(v);
            }
        }
        check(a.set_lists(a.ctx, static_cast<int>(views.size()), views.data(), nullptr, 0, numGroupPairs, 0, 1), "set_lists");
        a.searchCalls++;
    }
    const double t1 = now();
    a.secondsSearch += search ? t1 - t0 : 0.0;

    int flags = FEPB200_DO_SR; /* freeenergydispatch.cpp:169-184 */
    if (stepWork.computeForces)
    {
        flags |= FEPB200_DO_FORCE;
    }
    if (stepWork.computeVirial)
    {
        flags |= FEPB200_DO_SHIFTFORCE;
    }
    if (stepWork.computeEnergy)
    {
        flags |= FEPB200_DO_POTENTIAL;
    }
    const bool doForeign = nLambda > 0 && stepWork.computeDhdl && softCore; /* :236 */
    if (doForeign)
    {
        flags |= FEPB200_DO_FOREIGNLAMBDA;
    }
    std::vector<double> vc(numGroupPairs, 0.0), vv(numGroupPairs, 0.0), foreignE(nLambda + 1, 0.0),
            foreignDvdl(2 * (nLambda + 1), 0.0);
    double dvdl[2] = { 0.0, 0.0 };
    float* f      = stepWork.computeForces ? as_rvec_array(forceWithShiftForces->force().data())[0] : nullptr;
    float* fshift = (stepWork.computeForces && stepWork.computeVirial)
                            ? as_rvec_array(forceWithShiftForces->shiftForces().data())[0]
                            : nullptr;
    const float* x  = as_rvec_array(coords.unpaddedConstArrayRef().data())[0];
    const float* sv = as_rvec_array(shiftvec.data())[0];
    check(a.compute(a.ctx, x, sv, flags, f, fshift, vc.data(), vv.data(), dvdl, foreignE.data(), foreignDvdl.data()),
          "compute");

    if (stepWork.computeEnergy)
    {
        for (int g = 0; g < numGroupPairs; g++)
        {
            enerd->grpp.energyGroupPairTerms[NonBondedEnergyTerms::CoulombSR][g] += static_cast<real>(vc[g]);
            enerd->grpp.energyGroupPairTerms[NonBondedEnergyTerms::LJSR][g] += static_cast<real>(vv[g]);
        }
    }
    auto& dvdlDest = softCore ? enerd->dvdl_nonlin : enerd->dvdl_lin; /* :397-410 */
    dvdlDest[FreeEnergyPerturbationCouplingType::Coul] += dvdl[0];
    dvdlDest[FreeEnergyPerturbationCouplingType::Vdw] += dvdl[1];
    if (doForeign)
    {
        for (int i = 0; i <= nLambda; i++) /* :298-305 */
        {
            gmx::EnumerationArray<FreeEnergyPerturbationCouplingType, real> d = { 0 };
            d[FreeEnergyPerturbationCouplingType::Coul]                       = foreignDvdl[2 * i];
            d[FreeEnergyPerturbationCouplingType::Vdw]                        = foreignDvdl[2 * i + 1];
            enerd->foreignLambdaTerms.accumulate(i, foreignE[i], d);
        }
    }
    a.calls++;
    a.secondsStep += now() - t1;
}

} // namespace fepb200shim

#endif
