/*
 * fepb200_nb_gpu_shim.h -- binding of the cluster-pair kernel of libfepb200.so (include/fepb200_nb.h) inside the fork's GPU
 * route (`mdrun -nb gpu`), the counterpart of fepb200_gpu_shim.h for the NON-perturbed pairs.  Included by
 * src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp through nbnxm_gpu_nb_fepb200.patch; active when GMX_FEPB200_NB is set.
 *
 *   gpu_init_atomdata  (nbnxm_gpu_data_mgmt.cpp:873-988)  -> setAtoms(): masked types + nbfp as the reference holds them
 *   gpu_init_pairlist  (nbnxm_gpu_data_mgmt.cpp:667-759)  -> setList(): NbnxnPairlistGpu::sci / cjPacked / excl of the locality
 *   gpu_launch_kernel  (cuda/nbnxm_cuda.cu:738-750)       -> step(): instead of the launch of nbnxn_kernel_*_cuda
 *        fepb200_nb_launch_device_float_energies(adat->xq [charges masked in .w], adat->shiftVec, adat->f, adat->fShift,
 *                                                adat->eLJ, adat->eElec)   -- ONE kernel launch per step
 *     on the locality's nbnxm stream, inside the fork's nb_k GPU timer.  Coordinates, forces, shift forces and energies stay
 *     in the fork's device buffers: its copy-back and reduction (gpu_launch_cpyback, gpu_common.h:139-191) run unchanged, and
 *     with GMX_FEPB200 set as well the perturbed pairs are added into the same adat->f by fepb200_gpu_shim.h.
 * The kernel reads the fork's own device copy of the list (fepb200_nb_use_device_list), so the cluster pairs the fork's dynamic
 * pruning kernels clear there are skipped by it too.  Flavours the library does not cover (LJ switch functions, LJ-PME, combination
 * rules are irrelevant -- plain table look-up) fall through to the fork's kernel: step() returns false.
 */
#ifndef FEPB200_NB_GPU_SHIM_H
#define FEPB200_NB_GPU_SHIM_H

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "gromacs/utility/fatalerror.h"

#include "fepb200_nb.h"

namespace fepb200nbgpu
{

struct Api
{
    decltype(&fepb200_nb_create)                 create        = nullptr;
    decltype(&fepb200_nb_last_error)             last_error    = nullptr;
    decltype(&fepb200_nb_set_stream)             set_stream    = nullptr;
    decltype(&fepb200_nb_set_params)             set_params    = nullptr;
    decltype(&fepb200_nb_set_nbfp)               set_nbfp      = nullptr;
    decltype(&fepb200_nb_set_atoms)              set_atoms     = nullptr;
    decltype(&fepb200_nb_set_pairlist)           set_pairlist  = nullptr;
    decltype(&fepb200_nb_use_device_list)        use_device_list = nullptr;
    decltype(&fepb200_nb_launch_device)          launch_device = nullptr;
    decltype(&fepb200_nb_export_energies_device) export_energies_device = nullptr;
    decltype(&fepb200_nb_launch_device_float_energies) launch_device_float_energies = nullptr;
    bool                                         loaded = false;
};

struct Locality
{
    fepb200_nb*                 h = nullptr;
    void*                       stream = nullptr;
    const fepb200_nb_sci*       sci = nullptr;
    const fepb200_nb_cj_packed* cj = nullptr;
    const fepb200_nb_excl*      excl = nullptr;
    int                         nsci = 0, ncj = 0, nexcl = 0;
    bool                        listDirty = false, atomsDirty = false;
    long                        steps = 0, lists = 0;
};

struct State
{
    std::vector<int>   type;
    std::vector<float> nbfp, zeroCharge;
    int                ntype = 0;
    Locality           loc[2];
    bool               noted = false;
    ~State()
    {
        for (int i = 0; i < 2; i++)
        {
            if (loc[i].steps > 0)
            {
                std::fprintf(stderr, "fepb200 nb GPU route: locality %d: %ld steps, %ld list hand-overs\n", i, loc[i].steps,
                             loc[i].lists);
            }
        }
    }
};

inline bool enabled()
{
    static const bool on = std::getenv("GMX_FEPB200_NB") != nullptr;
    return on;
}

inline Api& api()
{
    static thread_local Api a;
    if (!a.loaded)
    {
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
        FEPB200_NB_SYM(set_stream);
        FEPB200_NB_SYM(set_params);
        FEPB200_NB_SYM(set_nbfp);
        FEPB200_NB_SYM(set_atoms);
        FEPB200_NB_SYM(set_pairlist);
        FEPB200_NB_SYM(use_device_list);
        FEPB200_NB_SYM(launch_device);
        FEPB200_NB_SYM(export_energies_device);
        FEPB200_NB_SYM(launch_device_float_energies);
#undef FEPB200_NB_SYM
        if (!a.create || !a.set_stream || !a.set_pairlist || !a.launch_device || !a.export_energies_device)
        {
            gmx_fatal(FARGS, "libfepb200.so lacks the fepb200_nb_* entry points the GPU route needs");
        }
    }
    return a;
}

/* said once when a run stays with the fork's kernel although GMX_FEPB200_NB is set */
inline void noteDeclined(const char* why)
{
    static thread_local bool said = false;
    if (!said)
    {
        std::fprintf(stderr, "NOTE: GMX_FEPB200_NB is set but the cluster pairs stay on the fork's kernel: %s\n", why);
        said = true;
    }
}

/* one state per NbnxmGpu (= per rank) */
inline State& state(const void* nb)
{
    static thread_local std::map<const void*, State> all;
    return all[nb];
}

inline void check(Locality& l, int rc, const char* what)
{
    if (rc != FEPB200_OK)
    {
        gmx_fatal(FARGS, "fepb200_nb %s failed (%d): %s", what, rc, api().last_error(l.h));
    }
}

/* gpu_init_atomdata: types (perturbed atoms masked by the reference already) and the LJ table, in nbat order */
inline void setAtoms(const void* nb, int natoms, const int* type, int ntype, const float* nbfp)
{
    if (!enabled())
    {
        return;
    }
    State& s = state(nb);
    s.type.assign(type, type + natoms);
    s.nbfp.assign(nbfp, nbfp + 2 * static_cast<size_t>(ntype) * ntype);
    s.zeroCharge.assign(natoms, 0.0F); /* the charges come from adat->xq.w (FEPB200_NB_Q_FROM_XQ) */
    s.ntype = ntype;
    s.loc[0].atomsDirty = s.loc[1].atomsDirty = true;
}

/* gpu_init_pairlist: the host list of the locality; it stays where it is until the next search */
inline void setList(const void* nb, int iloc, int nsci, const void* sci, int ncj, const void* cj, int nexcl, const void* excl)
{
    if (!enabled())
    {
        return;
    }
    Locality& l = state(nb).loc[iloc & 1];
    l.sci       = static_cast<const fepb200_nb_sci*>(sci);
    l.cj        = static_cast<const fepb200_nb_cj_packed*>(cj);
    l.excl      = static_cast<const fepb200_nb_excl*>(excl);
    l.nsci      = nsci;
    l.ncj       = ncj;
    l.nexcl     = nexcl;
    l.listDirty = true;
}

/* gpu_launch_kernel.  Returns false when this launch stays with the fork's kernel. */
inline bool step(const void* nb, int iloc, int device, void* stream, const fepb200_params& p, bool computeEnergy,
                 bool computeVirial, const float* d_xq, const float* d_shiftVec, float* d_f, float* d_fShift, float* d_eLJ,
                 float* d_eElec, const void* d_sci, const void* d_cjPacked, const void* d_excl)
{
    if (!enabled())
    {
        return false;
    }
    Api&      a = api();
    State&    s = state(nb);
    Locality& l = s.loc[iloc & 1];
    if (s.type.empty() || l.sci == nullptr)
    {
        return false;
    }
    if (!l.h)
    {
        if (a.create(&l.h, device) != FEPB200_OK)
        {
            gmx_fatal(FARGS, "fepb200_nb_create failed: %s", a.last_error(nullptr));
        }
        if (!s.noted)
        {
            std::fprintf(stderr, "NOTE: non-perturbed cluster pairs (GPU route) are computed by libfepb200 (fepb200_nb_*)\n");
            s.noted = true;
        }
    }
    if (l.stream != stream)
    {
        check(l, a.set_stream(l.h, stream), "set_stream");
        l.stream = stream;
    }
    check(l, a.set_params(l.h, &p), "set_params");
    if (l.atomsDirty)
    {
        check(l, a.set_nbfp(l.h, s.ntype, s.nbfp.data()), "set_nbfp");
        check(l, a.set_atoms(l.h, static_cast<int>(s.type.size()), s.type.data(), s.zeroCharge.data()), "set_atoms");
        l.atomsDirty = false;
        l.listDirty  = true; /* the list indexes these atoms */
    }
    if (l.listDirty)
    {
        check(l, a.set_pairlist(l.h, l.nsci, l.sci, l.ncj, l.cj, l.nexcl, l.excl), "set_pairlist");
        l.listDirty = false;
        l.lists++;
        /* the kernel reads the FORK's device copy of the list (same structure, uploaded by gpu_init_pairlist on this stream):
         * the i-cluster masks its dynamic pruning kernels clear are then skipped by our kernel as well.
         * GMX_FEPB200_NB_OWN_LIST keeps the library's own (unpruned) copy. */
        if (a.use_device_list && d_sci && d_cjPacked && d_excl && std::getenv("GMX_FEPB200_NB_OWN_LIST") == nullptr)
        {
            check(l, a.use_device_list(l.h, static_cast<const fepb200_nb_sci*>(d_sci),
                                       static_cast<const fepb200_nb_cj_packed*>(d_cjPacked),
                                       static_cast<const fepb200_nb_excl*>(d_excl)),
                  "use_device_list");
        }
    }
    int flags = FEPB200_DO_FORCE | FEPB200_NB_Q_FROM_XQ | FEPB200_NB_SHIFTVEC_ON_DEVICE;
    flags |= computeVirial ? FEPB200_DO_SHIFTFORCE : 0;
    flags |= computeEnergy ? FEPB200_DO_POTENTIAL : 0;
    if (a.launch_device_float_energies)
    {
        /* one kernel launch, energies added into the fork's float accumulators by the kernel itself */
        check(l, a.launch_device_float_energies(l.h, d_xq, d_shiftVec, flags, d_f, d_fShift, d_eLJ, d_eElec),
              "launch_device_float_energies");
    }
    else
    {
        check(l, a.launch_device(l.h, d_xq, d_shiftVec, flags, d_f, d_fShift, nullptr), "launch_device");
        if (computeEnergy)
        {
            check(l, a.export_energies_device(l.h, d_eLJ, d_eElec), "export_energies_device");
        }
    }
    l.steps++;
    return true;
}

} // namespace fepb200nbgpu

#endif
