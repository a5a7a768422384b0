/*
 * fepb200_gpu_shim.h -- libfepb200.so inside the fork's GPU route, `mdrun -nb gpu -fep gpu`.
 *
 * Included by the fork's src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp and src/gromacs/nbnxm/cuda/nbnxm_cuda.cu
 * through nbnxm_gpu_fepb200.patch.  With GMX_FEPB200 set, the launches of the fork's FEP kernels
 * (k_calc_nb_fep, k_calc_nb_fep_foreign; nbnxm_cuda.cu:755-851) are replaced by the library, device-resident:
 *
 *   hook                          fork function (file:line)                             what it hands over
 *   setInteractionConstants()     gpu_init                  (nbnxm_gpu_data_mgmt.cpp:538) interaction_const_t (read again every step:
 *                                                                                       PME tuning changes it in place)
 *   setLambdas()                  cuda_copy_fepparams       (:491-536)                  lambda_q, lambda_v, all_lambda at set-up
 *   setCurrentLambdas()           do_force                  (mdlib/sim_util.cpp:1764)   lambda_q, lambda_v of the coming step (the
 *                                                                                       fork uploads them once: SURVEY 2e-6)
 *   setAtoms()                    gpu_init_atomdata         (:990-1040)                 qA, qB, typeA, typeB and the nbfp table in
 *                                                                                       nbat (grid) order
 *   setShiftVectors()             gpu_upload_shiftvec       (:647-664)                  the 45 shift vectors
 *   setList()                     gpu_init_feppairlist      (:761-871)                  the FEP list remapped to nbat indices
 *   step()                        gpu_launch_kernel         (cuda/nbnxm_cuda.cu:755)    per force step, on the locality's stream
 *                                                                                       (called through Nbnxm::fepb200LaunchStep,
 *                                                                                       defined in nbnxm_gpu_data_mgmt.cpp):
 *       fepb200_gather_xq_device(adat->xq) -> fepb200_launch -> fepb200_add_forces_device(adat->f)
 *       -> fepb200_export_scalars_device(adat->eLJ, eElec, dvdlLJ, dvdlElec, e*Foreign, dvdl*Foreign, fShift)
 *
 * Coordinates, forces and scalars never leave the device; the fork's copy-back (gpu_launch_cpyback) and
 * reduction (gpu_common.h:139-191) run unchanged on what the library added into the fork's own buffers.
 * One library context per (NbnxmGpu, locality), on that locality's stream.
 */
#ifndef FEPB200_GPU_SHIM_H
#define FEPB200_GPU_SHIM_H

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "gromacs/mdtypes/interaction_const.h"

#include "fepb200_shim_common.h"

namespace fepb200gpu
{

using fepb200shim::Api;

struct LocalityState
{
    fepb200_ctx*     ctx = nullptr;
    void*            stream = nullptr;
    std::vector<int> iinr, shift, jindex, jjnr, excl, gid;
    std::vector<int> atomMap; /* index used by the list -> nbat index (empty: the list is in nbat indices already) */
    /* versions of the list / of the shared state this context holds (0 = nothing yet) */
    long   listVersion = 0, listApplied = 0, atomsApplied = 0, lambdasApplied = 0;
    bool   haveParams = false, lambdasSent = false;
    fepb200_params lastParams{};
    long   steps = 0, handovers = 0;
    double seconds = 0, secondsHandover = 0;
};

struct State
{
    const interaction_const_t* ic = nullptr;
    /* atoms in nbat order */
    std::vector<float> qA, qB, nbfp;
    std::vector<int>   typeA, typeB;
    int                numTypes     = 0;
    long               atomsVersion = 0;
    /* lambdas */
    std::vector<float> lambda = std::vector<float>(FEPB200_NUM_LAMBDA_COMPONENTS, 0.0F), allCoul, allVdw;
    long               lambdasVersion = 0;
    std::vector<float> shiftVec = std::vector<float>(3 * FEPB200_NUM_SHIFT_VECTORS, 0.0F);
    std::map<int, LocalityState> loc;
};

inline std::mutex& mutex()
{
    static std::mutex m;
    return m;
}

/* keyed by the NbnxmGpu object of the rank */
inline State& state(const void* nb)
{
    static std::map<const void*, State> all;
    std::lock_guard<std::mutex>         lock(mutex());
    return all[nb];
}

inline bool enabled()
{
    return fepb200shim::enabled();
}

inline void setInteractionConstants(const void* nb, const interaction_const_t* ic)
{
    if (enabled())
    {
        state(nb).ic = ic;
    }
}

inline void setLambdas(const void* nb, float lambdaCoul, float lambdaVdw, int nLambda, const float* allCoul, const float* allVdw)
{
    if (!enabled())
    {
        return;
    }
    State& s                        = state(nb);
    s.lambda[FEPB200_LAMBDA_COUL]   = lambdaCoul;
    s.lambda[FEPB200_LAMBDA_VDW]    = lambdaVdw;
    s.allCoul.assign(allCoul, allCoul + (nLambda > 0 ? nLambda : 0));
    s.allVdw.assign(allVdw, allVdw + (nLambda > 0 ? nLambda : 0));
    s.lambdasVersion++;
}

/* The lambdas of the coming step (hook in do_force, mdlib/sim_util.cpp, beside gpu_upload_shiftvec): closes the
 * fork's gap that lambda is uploaded once at set-up (nbnxm_setup.cpp:465-486), so that runs whose lambda moves
 * (delta-lambda / slow growth) see the current value, like the CPU route (freeenergydispatch.cpp:236-253). */
inline void setCurrentLambdas(const void* nb, float lambdaCoul, float lambdaVdw)
{
    if (!enabled())
    {
        return;
    }
    State& s = state(nb);
    if (s.lambda[FEPB200_LAMBDA_COUL] != lambdaCoul || s.lambda[FEPB200_LAMBDA_VDW] != lambdaVdw)
    {
        s.lambda[FEPB200_LAMBDA_COUL] = lambdaCoul;
        s.lambda[FEPB200_LAMBDA_VDW]  = lambdaVdw;
        s.lambdasVersion++;
    }
}

template<typename RealVector, typename IntVector>
inline void setAtoms(const void* nb, int numAtoms, const RealVector& qA, const RealVector& qB, const IntVector& typeA,
                     const IntVector& typeB, int numTypes, const RealVector& nbfp)
{
    if (!enabled())
    {
        return;
    }
    State& s = state(nb);
    s.qA.assign(qA.begin(), qA.begin() + numAtoms);
    s.qB.assign(qB.begin(), qB.begin() + numAtoms);
    s.typeA.assign(typeA.begin(), typeA.begin() + numAtoms);
    s.typeB.assign(typeB.begin(), typeB.begin() + numAtoms);
    s.numTypes = numTypes;
    s.nbfp.assign(nbfp.begin(), nbfp.begin() + 2 * static_cast<size_t>(numTypes) * numTypes);
    s.atomsVersion++;
}

inline void setShiftVectors(const void* nb, const float* shiftVec)
{
    if (enabled())
    {
        State& s = state(nb);
        s.shiftVec.assign(shiftVec, shiftVec + 3 * FEPB200_NUM_SHIFT_VECTORS);
    }
}

/* atomMap (may be NULL): the list's atom index -> nbat index, the fork's atomIndicesInv; the library applies it on the
 * device (fepb200_set_lists) instead of the host loops of gpu_init_feppairlist (nbnxm_gpu_data_mgmt.cpp:763-787) */
inline void setList(const void* nb, int iloc, int nri, const int* iinr, const int* shift, const int* jindex, int nrj,
                    const int* jjnr, const int* excl, const int* atomMap = nullptr, int nMap = 0)
{
    if (!enabled())
    {
        return;
    }
    LocalityState& l = state(nb).loc[iloc];
    l.iinr.assign(iinr, iinr + nri);
    l.shift.assign(shift, shift + nri);
    l.jindex.assign(jindex, jindex + nri + 1);
    l.jjnr.assign(jjnr, jjnr + nrj);
    l.excl.assign(excl, excl + nrj);
    l.gid.assign(nri, 0); /* the GPU route has one energy group */
    l.atomMap.assign(atomMap, atomMap + (atomMap ? nMap : 0));
    l.listVersion++;
}

inline void check(Api& a, fepb200_ctx* ctx, int rc, const char* what)
{
    if (rc != FEPB200_OK)
    {
        gmx_fatal(FARGS, "fepb200 (GPU route) %s failed (%d): %s", what, rc, a.last_error(ctx));
    }
}

/* One force step of one locality.  d_xq: float4[numAtoms]; d_f: float3[numAtoms]; the nine output buffers of
 * NBAtomDataGpu; device: the CUDA device ordinal of the rank; stream: the cudaStream_t of the locality. */
inline void step(const void* nb, int iloc, int device, void* stream, bool twoStreams, bool computeEnergy, bool computeVirial,
                 bool computeForeign, const float* d_xq, float* d_f, float* eLJ, float* eElec, float* dvdlLJ, float* dvdlElec, float* eLJForeign,
                 float* eElecForeign, float* dvdlLJForeign, float* dvdlElecForeign, float* fShift)
{
    fepb200shim::loadSymbols();
    Api&           a = fepb200shim::api();
    State&         s = state(nb);
    LocalityState& l = s.loc[iloc];
    const double   t0 = fepb200shim::now();
    if (!a.gather_xq_device || !a.launch || !a.add_forces_device || !a.export_scalars_device || !a.set_stream)
    {
        gmx_fatal(FARGS, "libfepb200.so lacks the device-resident entry points the GPU route needs");
    }
    if (!l.ctx)
    {
        const int rc = a.create(&l.ctx, device);
        if (rc != FEPB200_OK)
        {
            gmx_fatal(FARGS, "fepb200_create failed (%d): %s", rc, a.last_error(nullptr));
        }
        std::fprintf(stderr, "NOTE: perturbed non-bonded pairs (GPU route, locality %d) are computed by %s\n", iloc,
                     a.describe(l.ctx));
    }
    if (l.stream != stream)
    {
        check(a, l.ctx, a.set_stream(l.ctx, stream), "set_stream");
        l.stream = stream;
    }
    if (!s.ic || s.atomsVersion == 0 || l.listVersion == 0)
    {
        gmx_fatal(FARGS, "fepb200 GPU route: constants, atoms or pair list were never handed over");
    }
    /* hand over what changed since this context last saw it, in the order the library wants:
     * constants, type table + atoms, lambdas, list (new atoms invalidate the list the library holds) */
    const fepb200_params p = fepb200shim::toParams(*s.ic);
    if (!l.haveParams || std::memcmp(&p, &l.lastParams, sizeof(p)) != 0)
    {
        check(a, l.ctx, a.set_params(l.ctx, &p), "set_params");
        l.lastParams = p;
        l.haveParams = true;
    }
    const bool newAtoms = l.atomsApplied != s.atomsVersion;
    if (newAtoms)
    {
        check(a, l.ctx, a.set_nbfp(l.ctx, s.numTypes, s.nbfp.data(), nullptr), "set_nbfp");
        check(a, l.ctx,
              a.set_atoms(l.ctx, static_cast<int>(s.qA.size()), s.qA.data(), s.qB.data(), s.typeA.data(), s.typeB.data()),
              "set_atoms");
        l.atomsApplied = s.atomsVersion;
    }
    if (!l.lambdasSent || l.lambdasApplied != s.lambdasVersion)
    {
        check(a, l.ctx,
              a.set_lambdas(l.ctx, s.lambda.data(), static_cast<int>(s.allCoul.size()), s.allCoul.data(), s.allVdw.data()),
              "set_lambdas");
        l.lambdasApplied = s.lambdasVersion;
        l.lambdasSent    = true;
    }
    bool handover = newAtoms;
    if (newAtoms || l.listApplied != l.listVersion)
    {
        fepb200_list_view v;
        v.nri      = static_cast<int>(l.iinr.size());
        v.iinr     = l.iinr.data();
        v.gid      = l.gid.data();
        v.shift    = l.shift.data();
        v.jindex   = l.jindex.data();
        v.jjnr     = l.jjnr.data();
        v.excl_fep = l.excl.data();
        check(a, l.ctx,
              a.set_lists(l.ctx, 1, &v, l.atomMap.empty() ? nullptr : l.atomMap.data(), static_cast<int>(l.atomMap.size()), 1, 0, 1),
              "set_lists");
        l.listApplied = l.listVersion;
        handover      = true;
    }
    const double tHandover = fepb200shim::now();

    int flags = FEPB200_DO_SR | FEPB200_DO_FORCE; /* the fork's FEP kernels always compute forces */
    flags |= computeEnergy ? FEPB200_DO_POTENTIAL : 0;
    flags |= computeVirial ? FEPB200_DO_SHIFTFORCE : 0;
    flags |= computeForeign ? FEPB200_DO_FOREIGNLAMBDA : 0;
    check(a, l.ctx, a.gather_xq_device(l.ctx, d_xq, s.shiftVec.data()), "gather_xq_device");
    check(a, l.ctx, a.launch(l.ctx, flags, nullptr), "launch");
    /* a rank with two localities runs them on two streams, and the fork's kernels of the other locality add into the
     * same f / fShift / energy buffers with atomicAdd while ours are being added: then ours must be atomic as well */
    const int outFlags = twoStreams ? FEPB200_ATOMIC_OUTPUTS : 0;
    check(a, l.ctx, a.add_forces_device(l.ctx, d_f, outFlags), "add_forces_device");
    /* the fork clears and copies back eLJ / eElec / dvdl* only on steps that need them (gpu_clear_outputs,
     * nbnxm_gpu_data_mgmt.cpp:1049-1061; gpu_launch_cpyback :1196-1296): nothing may be added on the other steps */
    check(a, l.ctx,
          a.export_scalars_device(l.ctx, flags | outFlags, eLJ, eElec, computeEnergy ? dvdlLJ : nullptr,
                                  computeEnergy ? dvdlElec : nullptr, eLJForeign, eElecForeign, dvdlLJForeign, dvdlElecForeign,
                                  computeVirial ? fShift : nullptr),
          "export_scalars_device");
    l.steps++;
    l.seconds += fepb200shim::now() - tHandover;
    if (handover)
    {
        l.handovers++;
        l.secondsHandover += tHandover - t0;
    }
    if (l.steps == 20 || l.steps == 600 || l.steps == 5000)
    {
        std::fprintf(stderr,
                     "fepb200 GPU route: locality %d, %ld steps, %.1f us of host time per step to enqueue (gather, kernels, "
                     "force and scalar hand-off); %ld hand-overs of atoms / list, %.1f us each\n",
                     iloc, l.steps, 1e6 * l.seconds / l.steps, l.handovers, l.handovers ? 1e6 * l.secondsHandover / l.handovers : 0.0);
    }
}

} // namespace fepb200gpu

#endif
