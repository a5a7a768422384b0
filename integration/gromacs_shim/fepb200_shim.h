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

#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gromacs/utility/fatalerror.h"

#include "fepb200.h"

namespace fepb200shim
{

struct Api
{
    decltype(&fepb200_create)      create      = nullptr;
    decltype(&fepb200_last_error)  last_error  = nullptr;
    decltype(&fepb200_describe)    describe    = nullptr;
    decltype(&fepb200_set_params)  set_params  = nullptr;
    decltype(&fepb200_set_nbfp)    set_nbfp    = nullptr;
    decltype(&fepb200_set_atoms)   set_atoms   = nullptr;
    decltype(&fepb200_set_list)    set_list    = nullptr;
    decltype(&fepb200_set_lambdas) set_lambdas = nullptr;
    decltype(&fepb200_compute)     compute     = nullptr;
    fepb200_ctx*                   ctx         = nullptr;
    bool                           tried = false, ok = false;
    long                           calls = 0, searchCalls = 0;
    /* what the library holds, so that only changes are handed over between search steps */
    fepb200_params     lastParams{};
    std::vector<float> lastLambda, lastAllCoul, lastAllVdw;
    /* wall time spent in here, printed at exit (the same interval the "NB FEP" cycle counter sees) */
    double secondsSearch = 0, secondsStep = 0;
    ~Api(); /* prints the timing summary of this rank */
};

inline double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* One context per calling thread: with thread-MPI the ranks of a domain-decomposed run are threads
 * of one process, each with its own pair lists and local atom numbering. */
inline Api& api()
{
    static thread_local Api a;
    return a;
}

/* Device of the next context: GMX_FEPB200_DEVICES = number of GPUs to spread the ranks over
 * (rank k of the process gets device k mod that number; default 1 = everything on device 0). */
inline int nextDevice()
{
    static std::atomic<int> next{ 0 };
    const char*             e = std::getenv("GMX_FEPB200_DEVICES");
    const int               n = e ? std::atoi(e) : 1;
    return next.fetch_add(1) % (n > 0 ? n : 1);
}

inline Api::~Api()
{
    if (calls > 0)
    {
        std::fprintf(stderr,
                     "fepb200 shim: %ld calls (%ld with a new pair list); per call %.1f us for the step "
                     "(fepb200_compute + result routing), per new list %.1f us (set_atoms + set_list)\n",
                     calls, searchCalls, 1e6 * secondsStep / calls, searchCalls > 0 ? 1e6 * secondsSearch / searchCalls : 0.0);
    }
}

inline bool enabled()
{
    static const bool on = std::getenv("GMX_FEPB200") != nullptr;
    return on;
}

inline void check(int rc, const char* what)
{
    if (rc != FEPB200_OK)
    {
        gmx_fatal(FARGS, "fepb200 %s failed (%d): %s", what, rc, api().last_error(api().ctx));
    }
}

inline void load()
{
    Api& a = api();
    if (a.tried)
    {
        return;
    }
    a.tried          = true;
    const char* path = std::getenv("GMX_FEPB200_LIB");
    void*       h    = dlopen(path ? path : "libfepb200.so", RTLD_NOW | RTLD_LOCAL);
    if (!h)
    {
        gmx_fatal(FARGS, "GMX_FEPB200 is set but the library cannot be loaded: %s", dlerror());
    }
#define FEPB200_SYM(name) a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "fepb200_" #name))
    FEPB200_SYM(create);
    FEPB200_SYM(last_error);
    FEPB200_SYM(describe);
    FEPB200_SYM(set_params);
    FEPB200_SYM(set_nbfp);
    FEPB200_SYM(set_atoms);
    FEPB200_SYM(set_list);
    FEPB200_SYM(set_lambdas);
    FEPB200_SYM(compute);
#undef FEPB200_SYM
    if (!a.create || !a.compute || !a.set_list)
    {
        gmx_fatal(FARGS, "libfepb200.so does not export the expected symbols");
    }
    const int rc = a.create(&a.ctx, nextDevice());
    if (rc != FEPB200_OK)
    {
        gmx_fatal(FARGS, "fepb200_create failed (%d): %s", rc, a.last_error(nullptr));
    }
    std::fprintf(stderr, "NOTE: perturbed non-bonded pairs are computed by %s\n", a.describe(a.ctx));
    a.ok = true;
}

inline fepb200_params toParams(const interaction_const_t& ic)
{
    const auto&    sc = *ic.softCoreParameters;
    fepb200_params p{};
    p.eeltype                  = static_cast<int>(ic.eeltype);
    p.vdwtype                  = static_cast<int>(ic.vdwtype);
    p.vdw_modifier             = static_cast<int>(ic.vdw_modifier);
    p.epsfac                   = ic.epsfac;
    p.rcoulomb                 = ic.rcoulomb;
    p.rvdw                     = ic.rvdw;
    p.rvdw_switch              = ic.rvdw_switch;
    p.reactionFieldCoefficient = ic.reactionFieldCoefficient;
    p.reactionFieldShift       = ic.reactionFieldShift;
    p.sh_ewald                 = ic.sh_ewald;
    p.sh_lj_ewald              = ic.sh_lj_ewald;
    p.ewaldcoeff_q             = ic.ewaldcoeff_q;
    p.ewaldcoeff_lj            = ic.ewaldcoeff_lj;
    p.dispersion_shift_cpot    = ic.dispersion_shift.cpot;
    p.repulsion_shift_cpot     = ic.repulsion_shift.cpot;
    p.softcoreType             = static_cast<int>(sc.softcoreType);
    p.alphaVdw                 = sc.alphaVdw;
    p.alphaCoulomb             = sc.alphaCoulomb;
    p.lambdaPower              = sc.lambdaPower;
    p.sigma6WithInvalidSigma   = sc.sigma6WithInvalidSigma;
    p.sigma6Minimum            = sc.sigma6Minimum;
    p.gapsysScaleLinpointVdW   = sc.gapsysScaleLinpointVdW;
    p.gapsysScaleLinpointCoul  = sc.gapsysScaleLinpointCoul;
    p.gapsysSigma6VdW          = sc.gapsysSigma6VdW;
    return p;
}

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
    load();
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
        /* the FEP pair lists of all localities and threads, concatenated (what the fork's
         * combine_fep_lists does for its GPU path, pairlist.cpp:2867) */
        std::vector<int> iinr, gid, shift, jindex(1, 0), jjnr, excl;
        const int        numLocalities = (pairlistSets.params().haveMultipleDomains ? 2 : 1);
        for (int l = 0; l < numLocalities; l++)
        {
            const auto lists = pairlistSets.pairlistSet(static_cast<gmx::InteractionLocality>(l)).fepLists();
            for (const auto& nl : lists)
            {
                for (int n = 0; n < nl->nri; n++)
                {
                    iinr.push_back(nl->iinr[n]);
                    gid.push_back(nl->gid[n]);
                    shift.push_back(nl->shift[n]);
                    for (int k = nl->jindex[n]; k < nl->jindex[n + 1]; k++)
                    {
                        jjnr.push_back(nl->jjnr[k]);
                        excl.push_back(nl->excl_fep[k]);
                    }
                    jindex.push_back(static_cast<int>(jjnr.size()));
                }
            }
        }
        check(a.set_list(a.ctx, static_cast<int>(iinr.size()), iinr.data(), gid.data(), shift.data(), jindex.data(),
                         jjnr.data(), excl.data(), numGroupPairs, 0, 1),
              "set_list");
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
