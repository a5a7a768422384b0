/*
 * free_energy_dispatch_gpu.h -- C++ host side of the drop-in: a class with the call surface of the
 * reference's FreeEnergyDispatch (src/gromacs/nbnxm/freeenergydispatch.h:63-110, .cpp:63-459) that
 * drives libfepb200.so through the C-ABI of include/fepb200.h instead of calling
 * gmx_nb_free_energy_kernel per OpenMP thread.
 *
 * The structs below are plain views with the fields the reference reads on this path; inside
 * GROMACS they are filled from t_nblist, interaction_const_t, t_lambda, StepWorkload and
 * gmx_enerdata_t (INTEGRATION.md shows the mapping).  No CUDA or torch types appear here.
 */
#ifndef FEPB200_FREE_ENERGY_DISPATCH_GPU_H
#define FEPB200_FREE_ENERGY_DISPATCH_GPU_H

#include <array>
#include <stdexcept>
#include <string>
#include <vector>

#include "fepb200.h"

namespace fepb200
{

/* mdtypes/nblist.h:41-55 */
struct NbListView
{
    int        nri      = 0;
    const int* iinr     = nullptr;
    const int* gid      = nullptr;
    const int* shift    = nullptr;
    const int* jindex   = nullptr; /* nri + 1 */
    const int* jjnr     = nullptr;
    const int* excl_fep = nullptr; /* may be null: all pairs included */
};

/* the members of gmx::StepWorkload the dispatcher looks at (freeenergydispatch.cpp:169-184,236) */
struct StepWork
{
    bool computeForces = true;
    bool computeVirial = false;
    bool computeEnergy = false;
    bool computeDhdl   = false;
};

/* t_lambda: the parts used here (inputrec.h:114-166) */
struct LambdaTable
{
    int                n_lambda = 0;
    std::vector<float> all_lambda_coul; /* all_lambda[FreeEnergyPerturbationCouplingType::Coul] */
    std::vector<float> all_lambda_vdw;  /* all_lambda[FreeEnergyPerturbationCouplingType::Vdw]  */
};

/* the parts of gmx_enerdata_t this path adds to (mdtypes/enerdata.h:78-205) */
struct EnergyData
{
    std::vector<double> vCoulombSR, vLJSR; /* grpp.energyGroupPairTerms[CoulombSR|LJSR], size G */
    std::array<double, FEPB200_NUM_LAMBDA_COMPONENTS> dvdl_lin{}, dvdl_nonlin{};
    /* ForeignLambdaTerms::accumulate(i, energy, dvdl): energies_[i] += energy, dhdl_[i] += sum dvdl */
    std::vector<double> foreignEnergies, foreignDhdl; /* size n_lambda + 1 */
    explicit EnergyData(int numEnergyGroupPairs = 1, int numLambdas = 0) :
        vCoulombSR(numEnergyGroupPairs, 0.0),
        vLJSR(numEnergyGroupPairs, 0.0),
        foreignEnergies(numLambdas + 1, 0.0),
        foreignDhdl(numLambdas + 1, 0.0)
    {
    }
};

class Error : public std::runtime_error
{
public:
    Error(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
    int code() const { return code_; }

private:
    int code_;
};

class FreeEnergyDispatchGpu
{
public:
    /* replaces FreeEnergyDispatch(numEnergyGroups) + Nbnxm::gpu_init(..., bFEP, n_lambda) */
    FreeEnergyDispatchGpu(int numEnergyGroups, int deviceId);
    ~FreeEnergyDispatchGpu();
    FreeEnergyDispatchGpu(const FreeEnergyDispatchGpu&) = delete;
    FreeEnergyDispatchGpu& operator=(const FreeEnergyDispatchGpu&) = delete;

    /* init time: interaction_const_t + SoftCoreParameters, nbfp tables (forcerec) */
    void setInteractionConstants(const fepb200_params& ic);
    void setNonbondedParameters(int ntype, const float* nbfp, const float* nbfpGrid);
    /* search steps: nbv->setAtomPropertiesAB + constructPairlist/gpu_init_feppairlist;
     * rank/numRanks: this process' share when one process per GPU splits the list */
    void setAtomPropertiesAB(int numAtoms, const float* chargeA, const float* chargeB, const int* typeA,
                             const int* typeB);
    void setPairlist(const NbListView& list, int rank = 0, int numRanks = 1);
    /* lambda[7] of the current state and the foreign-lambda table (t_lambda) */
    void setLambdas(const float* lambda, const LambdaTable& fepvals);

    /* Same role and argument meaning as FreeEnergyDispatch::dispatchFreeEnergyKernels
     * (freeenergydispatch.cpp:312-413): forces and shift forces are ADDED to the caller's rvec
     * arrays, energies to enerd->grpp, dV/dlambda to dvdl_nonlin (soft-core active) or dvdl_lin,
     * and on dhdl steps with soft-core the L+1 foreign energies are accumulated. */
    void dispatchFreeEnergyKernels(const float* coords /* rvec[numAtoms] */, const float* shiftVectors /* rvec[45] */,
                                   const StepWork& stepWork, float* forces /* rvec[numAtoms] */,
                                   float* shiftForces /* rvec[45] */, EnergyData* enerd);

    bool        haveSoftCore() const;
    int         numEnergyGroupPairs() const { return numGroupPairs_; }
    std::string describe() const;
    fepb200_ctx* context() { return ctx_; }

private:
    void check(int rc) const;

    fepb200_ctx*   ctx_           = nullptr;
    int            numGroupPairs_ = 1;
    int            numLambdas_    = 0;
    fepb200_params params_{};
    bool           haveParams_ = false;
    std::vector<double> vc_, vv_, foreignE_, foreignDvdl_;
};

} // namespace fepb200

#endif
