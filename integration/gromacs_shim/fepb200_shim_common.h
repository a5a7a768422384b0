/*
 * fepb200_shim_common.h -- what both reference-side bindings of libfepb200.so share: the dlopen()ed entry
 * points, one library context per calling thread for the CPU route (fepb200_shim.h), the translation of
 * interaction_const_t into fepb200_params.  Included by fepb200_shim.h (hook in
 * src/gromacs/nbnxm/freeenergydispatch.cpp), fepb200_gpu_shim.h (hooks in the fork's GPU route) and
 * fepb200_pairs14_shim.h (hook in src/gromacs/listed_forces/pairs.cpp).
 */
#ifndef FEPB200_SHIM_COMMON_H
#define FEPB200_SHIM_COMMON_H

#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
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
    decltype(&fepb200_set_lists)   set_lists   = nullptr;
    decltype(&fepb200_set_lambdas) set_lambdas = nullptr;
    decltype(&fepb200_compute)     compute     = nullptr;
    /* device-resident entry points (GPU route, fepb200_gpu_shim.h) */
    decltype(&fepb200_set_stream)            set_stream            = nullptr;
    decltype(&fepb200_gather_xq_device)      gather_xq_device      = nullptr;
    decltype(&fepb200_launch)                launch                = nullptr;
    decltype(&fepb200_add_forces_device)     add_forces_device     = nullptr;
    decltype(&fepb200_export_scalars_device) export_scalars_device = nullptr;
    /* perturbed 1-4 pairs (fepb200_pairs14_shim.h) */
    decltype(&fepb200_pairs14_create)     pairs14_create     = nullptr;
    decltype(&fepb200_pairs14_last_error) pairs14_last_error = nullptr;
    decltype(&fepb200_pairs14_set_params) pairs14_set_params = nullptr;
    decltype(&fepb200_pairs14_set_pairs)  pairs14_set_pairs  = nullptr;
    decltype(&fepb200_pairs14_compute)    pairs14_compute    = nullptr;
    decltype(&fepb200_pairs14_compute_foreign) pairs14_compute_foreign = nullptr;
    fepb200_ctx*                   ctx         = nullptr;
    bool                           symbols = false, tried = false, ok = false;
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

/* Device of a rank's contexts.  GMX_FEPB200_DEVICES = number of GPUs to spread the ranks over (default 1 = everything on
 * device 0): the k-th rank to ask gets device k mod that number, and EVERY context of that rank -- the non-bonded one and the
 * per-OpenMP-thread handles of the perturbed 1-4 pairs -- gets the same one.  A rank is identified by the address of its
 * interaction_const_t (one per t_forcerec, i.e. per thread-MPI rank; all its OpenMP threads see the same object).  The GPU
 * route does not come here: it passes the device GROMACS assigned to the rank (fepb200_gpu_shim.h). */
inline int deviceForRank(const void* rankKey)
{
    static std::mutex                 mutex;
    static std::map<const void*, int> assigned;
    std::lock_guard<std::mutex>       guard(mutex);
    auto                              it = assigned.find(rankKey);
    if (it == assigned.end())
    {
        const char* e = std::getenv("GMX_FEPB200_DEVICES");
        const int   n = e ? std::atoi(e) : 1;
        it            = assigned.emplace(rankKey, static_cast<int>(assigned.size()) % (n > 0 ? n : 1)).first;
    }
    return it->second;
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

/* dlopen() + the entry points, once per calling thread */
inline void loadSymbols()
{
    Api& a = api();
    if (a.symbols)
    {
        return;
    }
    a.symbols        = true;
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
    FEPB200_SYM(set_lists);
    FEPB200_SYM(set_lambdas);
    FEPB200_SYM(compute);
    FEPB200_SYM(set_stream);
    FEPB200_SYM(gather_xq_device);
    FEPB200_SYM(launch);
    FEPB200_SYM(add_forces_device);
    FEPB200_SYM(export_scalars_device);
    FEPB200_SYM(pairs14_create);
    FEPB200_SYM(pairs14_last_error);
    FEPB200_SYM(pairs14_set_params);
    FEPB200_SYM(pairs14_set_pairs);
    FEPB200_SYM(pairs14_compute);
    FEPB200_SYM(pairs14_compute_foreign);
#undef FEPB200_SYM
    if (!a.create || !a.compute || !a.set_list || !a.set_lists)
    {
        gmx_fatal(FARGS, "libfepb200.so does not export the expected symbols");
    }
}

/* ... and the context of the CPU route of this thread */
inline void load(const void* rankKey)
{
    Api& a = api();
    if (a.tried)
    {
        return;
    }
    a.tried = true;
    loadSymbols();
    const int rc = a.create(&a.ctx, deviceForRank(rankKey));
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

} // namespace fepb200shim

#endif
