/*
 * fepb200_nb_shim.h -- reference-side binding of the non-perturbed cluster-pair kernel of libfepb200.so
 * (include/fepb200_nb.h).  Included by src/gromacs/nbnxm/kerneldispatch.cpp through
 * kerneldispatch_fepb200.patch: where nonbonded_verlet_t::dispatchNonbondedKernel calls the reference kernel for
 * GPU-layout pair lists (nbnxn_kernel_gpu_ref, kerneldispatch.cpp:479-490, the route `GMX_EMULATE_GPU=1` selects),
 * the call goes to the B200 instead when GMX_FEPB200_NB is set.  Everything is handed over as the reference holds it:
 *   NbnxnPairlistGpu::sci / cjPacked.list_ / excl       -> fepb200_nb_set_pairlist   (search steps)
 *   nbat->params().type, charges in nbat->x()[.w], nbfp -> fepb200_nb_set_nbfp / set_atoms (search steps; the reference
 *                                                          has masked the perturbed atoms already, atomdata.cpp:930-964)
 *   nbat->x() (nbatXYZQ), shift vectors                 -> fepb200_nb_compute_xyzq   (every step)
 *   nbat->out[0].f / fshift, Vc[0], Vvdw[0]             <- accumulated, or overwritten when clearF says so
 * dlopen()ed like the other bindings: the GROMACS build system is untouched.
 */
#ifndef FEPB200_NB_SHIM_H
#define FEPB200_NB_SHIM_H

#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/simulation_workload.h"
#include "gromacs/nbnxm/atomdata.h"
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/pairlist.h"
#include "gromacs/utility/fatalerror.h"

#include "fepb200_nb.h"

namespace fepb200nbshim
{

struct Api
{
    decltype(&fepb200_nb_create)       create       = nullptr;
    decltype(&fepb200_nb_last_error)   last_error   = nullptr;
    decltype(&fepb200_nb_set_params)   set_params   = nullptr;
    decltype(&fepb200_nb_set_nbfp)     set_nbfp     = nullptr;
    decltype(&fepb200_nb_set_atoms)    set_atoms    = nullptr;
    decltype(&fepb200_nb_set_pairlist) set_pairlist = nullptr;
    decltype(&fepb200_nb_compute_xyzq) compute_xyzq = nullptr;
    fepb200_nb*                        h            = nullptr;
    bool                               loaded       = false;
    long                               calls = 0, lists = 0;
    double                             secondsStep = 0, secondsList = 0;
    std::vector<float>                 q, f;
    ~Api()
    {
        if (calls > 0)
        {
            std::fprintf(stderr,
                         "fepb200 nb shim: %ld calls (%ld with a new pair list); per call %.1f us for the step, per new list "
                         "%.1f us\n",
                         calls, lists, 1e6 * secondsStep / calls, lists > 0 ? 1e6 * secondsList / lists : 0.0);
        }
    }
};

/* thread-MPI ranks are threads: one handle each, and one per interaction locality (each has its own pair list) */
inline Api& api(int locality = -1)
{
    static thread_local Api a[2];
    static thread_local int current = 0;
    if (locality >= 0)
    {
        current = locality & 1;
    }
    return a[current];
}

inline bool enabled()
{
    static const bool on = std::getenv("GMX_FEPB200_NB") != nullptr;
    return on;
}

inline double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline void check(int rc, const char* what)
{
    if (rc != FEPB200_OK)
    {
        gmx_fatal(FARGS, "fepb200_nb %s failed (%d): %s", what, rc, api().last_error(api().h));
    }
}

inline void load()
{
    Api& a = api();
    if (a.loaded)
    {
        return;
    }
    a.loaded         = true;
    const char* path = std::getenv("GMX_FEPB200_LIB");
    void*       lib  = dlopen(path ? path : "libfepb200.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib)
    {
        gmx_fatal(FARGS, "GMX_FEPB200_NB is set but the library cannot be loaded: %s", dlerror());
    }
#define FEPB200_NB_SYM(name) a.name = reinterpret_cast<decltype(a.name)>(dlsym(lib, "fepb200_nb_" #name))
    FEPB200_NB_SYM(create);
    FEPB200_NB_SYM(last_error);
    FEPB200_NB_SYM(set_params);
    FEPB200_NB_SYM(set_nbfp);
    FEPB200_NB_SYM(set_atoms);
    FEPB200_NB_SYM(set_pairlist);
    FEPB200_NB_SYM(compute_xyzq);
#undef FEPB200_NB_SYM
    if (!a.create || !a.last_error || !a.set_params || !a.set_nbfp || !a.set_atoms || !a.set_pairlist || !a.compute_xyzq)
    {
        gmx_fatal(FARGS, "libfepb200.so lacks the fepb200_nb_* entry points");
    }
    const char* dev = std::getenv("GMX_FEPB200_DEVICE");
    if (a.create(&a.h, dev ? std::atoi(dev) : 0) != FEPB200_OK)
    {
        gmx_fatal(FARGS, "fepb200_nb_create failed: %s", a.last_error(nullptr));
    }
}

/* Same arguments as nbnxn_kernel_gpu_ref (kernels_reference/kernel_gpu_ref.h).  Returns false when the binding is off. */
inline bool dispatch(int                            locality,
                     const NbnxnPairlistGpu*        nbl,
                     const nbnxn_atomdata_t*        nbat,
                     const interaction_const_t*     ic,
                     gmx::ArrayRef<const gmx::RVec> shiftvec,
                     const gmx::StepWorkload&       stepWork,
                     int                            clearF,
                     gmx::ArrayRef<real>            f,
                     real*                          fshift,
                     real*                          Vc,
                     real*                          Vvdw)
{
    if (!enabled())
    {
        return false;
    }
    static_assert(sizeof(real) == sizeof(float), "the binding is for the mixed-precision build");
    static_assert(sizeof(nbnxn_sci_t) == sizeof(fepb200_nb_sci) && sizeof(nbnxn_cj_packed_t) == sizeof(fepb200_nb_cj_packed)
                          && sizeof(nbnxn_excl_t) == sizeof(fepb200_nb_excl),
                  "list structures must be the reference's, byte for byte");
    Api& a = api(locality);
    load();
    const double t0     = now();
    const int    natoms = nbat->numAtoms();
    if (nbat->XFormat != nbatXYZQ || nbat->xstride != STRIDE_XYZQ || nbat->fstride != DIM)
    {
        gmx_fatal(FARGS, "fepb200_nb: the atom data is not in the GPU layout (xyzq / xyz)");
    }
    if (stepWork.doNeighborSearch || a.lists == 0)
    {
        fepb200_params p{};
        p.eeltype                  = static_cast<int>(ic->eeltype);
        p.epsfac                   = ic->epsfac;
        p.rcoulomb                 = ic->rcoulomb;
        p.rvdw                     = ic->rvdw;
        p.reactionFieldCoefficient = ic->reactionFieldCoefficient;
        p.reactionFieldShift       = ic->reactionFieldShift;
        p.sh_ewald                 = ic->sh_ewald;
        p.ewaldcoeff_q             = ic->ewaldcoeff_q;
        p.dispersion_shift_cpot    = ic->dispersion_shift.cpot;
        p.repulsion_shift_cpot     = ic->repulsion_shift.cpot;
        check(a.set_params(a.h, &p), "set_params");
        check(a.set_nbfp(a.h, nbat->params().numTypes, nbat->params().nbfp.data()), "set_nbfp");
        /* the reference masked the perturbed atoms already: types in params().type, charges in the .w of x() */
        a.q.resize(natoms);
        for (int i = 0; i < natoms; i++)
        {
            a.q[i] = nbat->x()[i * STRIDE_XYZQ + 3];
        }
        check(a.set_atoms(a.h, natoms, nbat->params().type.data(), a.q.data()), "set_atoms");
        check(a.set_pairlist(a.h, static_cast<int>(nbl->sci.size()), reinterpret_cast<const fepb200_nb_sci*>(nbl->sci.data()),
                             static_cast<int>(nbl->cjPacked.list_.size()),
                             reinterpret_cast<const fepb200_nb_cj_packed*>(nbl->cjPacked.list_.data()),
                             static_cast<int>(nbl->excl.size()), reinterpret_cast<const fepb200_nb_excl*>(nbl->excl.data())),
              "set_pairlist");
        a.lists++;
        a.secondsList += now() - t0;
    }
    const double t1    = now();
    int          flags = FEPB200_DO_FORCE | FEPB200_NB_Q_FROM_XQ;
    if (stepWork.computeVirial)
    {
        flags |= FEPB200_DO_SHIFTFORCE;
    }
    if (stepWork.computeEnergy)
    {
        flags |= FEPB200_DO_POTENTIAL;
    }
    if (clearF == enbvClearFYes)
    {
        /* the reference clears f only (kernel_gpu_ref.cpp:80-86); fshift and the energies are always added to */
        std::fill(f.begin(), f.end(), 0.0_real);
    }
    float  fshiftDummy[3 * gmx::c_numShiftVectors];
    double vc = 0, vvdw = 0;
    check(a.compute_xyzq(a.h, nbat->x().data(), reinterpret_cast<const float*>(shiftvec.data()), flags, f.data(),
                         stepWork.computeVirial ? fshift : fshiftDummy, &vc, &vvdw),
          "compute_xyzq");
    if (stepWork.computeEnergy)
    {
        Vc[0] += static_cast<real>(vc);
        Vvdw[0] += static_cast<real>(vvdw);
    }
    a.calls++;
    a.secondsStep += now() - t1;
    return true;
}

} // namespace fepb200nbshim

#endif
