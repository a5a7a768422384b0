/*
 * free_energy_dispatch_gpu.cpp -- see free_energy_dispatch_gpu.h.  Flag assembly, the condition for
 * the foreign-lambda evaluation and the routing of the results follow
 * src/gromacs/nbnxm/freeenergydispatch.cpp:139-145 (haveSoftCore), :169-184 (flags), :236 (foreign
 * condition), :298-305 (ForeignLambdaTerms::accumulate), :395-410 (dvdl_lin / dvdl_nonlin).
 */
#include "free_energy_dispatch_gpu.h"

namespace fepb200
{

FreeEnergyDispatchGpu::FreeEnergyDispatchGpu(int numEnergyGroups, int deviceId) :
    numGroupPairs_(numEnergyGroups * numEnergyGroups)
{
    const int rc = fepb200_create(&ctx_, deviceId);
    if (rc != FEPB200_OK)
    {
        throw Error(rc, std::string("fepb200_create: ") + fepb200_last_error(nullptr));
    }
}

FreeEnergyDispatchGpu::~FreeEnergyDispatchGpu()
{
    fepb200_destroy(ctx_);
}

void FreeEnergyDispatchGpu::check(int rc) const
{
    if (rc != FEPB200_OK)
    {
        throw Error(rc, fepb200_last_error(ctx_));
    }
}

void FreeEnergyDispatchGpu::setInteractionConstants(const fepb200_params& ic)
{
    check(fepb200_set_params(ctx_, &ic));
    params_     = ic;
    haveParams_ = true;
}

void FreeEnergyDispatchGpu::setNonbondedParameters(int ntype, const float* nbfp, const float* nbfpGrid)
{
    check(fepb200_set_nbfp(ctx_, ntype, nbfp, nbfpGrid));
}

void FreeEnergyDispatchGpu::setAtomPropertiesAB(int numAtoms, const float* chargeA, const float* chargeB,
                                                const int* typeA, const int* typeB)
{
    check(fepb200_set_atoms(ctx_, numAtoms, chargeA, chargeB, typeA, typeB));
}

void FreeEnergyDispatchGpu::setPairlist(const NbListView& l, int rank, int numRanks)
{
    check(fepb200_set_list(ctx_, l.nri, l.iinr, l.gid, l.shift, l.jindex, l.jjnr, l.excl_fep, numGroupPairs_, rank,
                           numRanks));
}

void FreeEnergyDispatchGpu::setLambdas(const float* lambda, const LambdaTable& fepvals)
{
    numLambdas_ = fepvals.n_lambda;
    check(fepb200_set_lambdas(ctx_, lambda, fepvals.n_lambda, fepvals.all_lambda_coul.data(),
                              fepvals.all_lambda_vdw.data()));
}

bool FreeEnergyDispatchGpu::haveSoftCore() const
{
    /* freeenergydispatch.cpp:139-145 */
    return (params_.softcoreType == FEPB200_SC_BEUTLER && (params_.alphaCoulomb != 0 || params_.alphaVdw != 0))
           || (params_.softcoreType == FEPB200_SC_GAPSYS
               && (params_.gapsysScaleLinpointCoul != 0 || params_.gapsysScaleLinpointVdW != 0));
}

std::string FreeEnergyDispatchGpu::describe() const
{
    return fepb200_describe(ctx_);
}

void FreeEnergyDispatchGpu::dispatchFreeEnergyKernels(const float* coords, const float* shiftVectors,
                                                      const StepWork& stepWork, float* forces, float* shiftForces,
                                                      EnergyData* enerd)
{
    int flags = FEPB200_DO_SR;
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
    const bool doForeign = numLambdas_ > 0 && stepWork.computeDhdl && haveSoftCore();
    if (doForeign)
    {
        flags |= FEPB200_DO_FOREIGNLAMBDA;
    }
    vc_.assign(numGroupPairs_, 0.0);
    vv_.assign(numGroupPairs_, 0.0);
    foreignE_.assign(numLambdas_ + 1, 0.0);
    foreignDvdl_.assign(2 * (numLambdas_ + 1), 0.0);
    double dvdl[2] = { 0.0, 0.0 };

    check(fepb200_compute(ctx_, coords, shiftVectors, flags, forces, shiftForces, vc_.data(), vv_.data(), dvdl,
                          foreignE_.data(), foreignDvdl_.data()));

    if (stepWork.computeEnergy)
    {
        for (int g = 0; g < numGroupPairs_; g++)
        {
            enerd->vCoulombSR[g] += vc_[g];
            enerd->vLJSR[g] += vv_[g];
        }
    }
    auto& dvdlDest = haveSoftCore() ? enerd->dvdl_nonlin : enerd->dvdl_lin;
    dvdlDest[FEPB200_LAMBDA_COUL] += dvdl[0];
    dvdlDest[FEPB200_LAMBDA_VDW] += dvdl[1];
    if (doForeign)
    {
        for (int i = 0; i <= numLambdas_; i++)
        {
            /* ForeignLambdaTerms::accumulate (mdtypes/enerdata.h:120-136) */
            enerd->foreignEnergies[i] += foreignE_[i];
            enerd->foreignDhdl[i] += foreignDvdl_[2 * i] + foreignDvdl_[2 * i + 1];
        }
    }
}

} // namespace fepb200
