/*
 * oracle/ref_build/nb_harness.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" driver around the UNMODIFIED reference kernel for GPU-layout cluster pair lists
 *   nbnxn_kernel_gpu_ref()   /root/reference/src/gromacs/nbnxm/kernels_reference/kernel_gpu_ref.cpp:54-354
 * which the Makefile next to this file compiles from where it lies under /root/reference (nothing
 * from the reference is copied into this repository).  The result, oracle/_ref/libnbref_{dp,sp}.so,
 * pins oracle/nb_oracle.c (tests/test_oracle_nb.py); only tests/ and bench.py's CPU legs load it.
 *
 * The reference's containers (NbnxnPairlistGpu, nbnxn_atomdata_t) have constructors that live in
 * translation units we do not compile (pairlist.cpp, atomdata.cpp).  The kernel only reads a handful
 * of their members, so the driver takes zeroed storage of the right size and constructs exactly those
 * members in place; the objects are never destroyed through their own destructors.
 */
#include "config.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <memory>
#include <new>
#include <vector>

#include "gromacs/gpu_utils/hostallocator.h"
#include "gromacs/math/vectypes.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/md_enums.h"
#include "gromacs/mdtypes/simulation_workload.h"
#define private public /* x_ and params_ of nbnxn_atomdata_t: filled member by member, see above */
#include "gromacs/nbnxm/atomdata.h"
#undef private
#include "gromacs/nbnxm/nbnxm.h"
#include "gromacs/nbnxm/pairlist.h"
#include "gromacs/nbnxm/kernels_reference/kernel_gpu_ref.h"
#include "gromacs/pbcutil/ishift.h"
#include "gromacs/utility/alignedallocator.h"
#include "gromacs/utility/arrayref.h"
#include "gromacs/utility/real.h"

/* ---- the symbols kernel_gpu_ref.o and the containers need from the rest of libgromacs ---- */
FILE* debug = nullptr;
[[noreturn]] void gmx_fatal(int /*fatal_errno*/, const std::filesystem::path& file, int line, const char* fmt, ...)
{
    std::va_list ap;
    va_start(ap, fmt);
    std::fprintf(stderr, "reference gmx_fatal at %s:%d: ", file.c_str(), line);
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
    std::abort();
}
namespace gmx
{
HostAllocationPolicy::HostAllocationPolicy(PinningPolicy policy) : pinningPolicy_(policy) {}
std::size_t HostAllocationPolicy::alignment() const noexcept
{
    return 128;
}
void* HostAllocationPolicy::malloc(std::size_t bytes) const noexcept
{
    void* p = nullptr;
    return posix_memalign(&p, 128, bytes ? bytes : 128) == 0 ? p : nullptr;
}
void HostAllocationPolicy::free(void* buffer) const noexcept
{
    std::free(buffer);
}
void* AlignedAllocationPolicy::malloc(std::size_t bytes)
{
    void* p = nullptr;
    return posix_memalign(&p, 128, bytes ? bytes : 128) == 0 ? p : nullptr;
}
void AlignedAllocationPolicy::free(void* p)
{
    std::free(p);
}
std::size_t AlignedAllocationPolicy::alignment()
{
    return 128;
}
namespace internal
{
[[noreturn]] void assertHandler(const char* condition, const char* msg, const char* func,
                                const std::filesystem::path& file, int line)
{
    std::fprintf(stderr, "reference assertion failed: %s (%s) in %s at %s:%d\n", condition, msg, func,
                 file.c_str(), line);
    std::abort();
}
} // namespace internal
} // namespace gmx

/* Constants of the interaction the kernel reads (kernel_gpu_ref.cpp:88-101,230-262) and the Ewald force
 * table it interpolates (:239-246); doubles at the interface, converted to the build's `real`. */
extern "C" struct nbref_params
{
    int    eeltype; /* CoulombInteractionType as int */
    double epsfac, rcoulomb, rvdw, rlist;
    double reactionFieldCoefficient, reactionFieldShift;
    double sh_ewald, ewaldcoeff_q;
    double dispersion_shift_cpot, repulsion_shift_cpot;
    double        tab_scale; /* points per nm */
    int           tab_size;
    const double* tableF;
};

namespace
{
template<typename T>
T* zeroedStorage()
{
    void* p = nullptr;
    if (posix_memalign(&p, 128, sizeof(T)) != 0)
    {
        std::abort();
    }
    std::memset(p, 0, sizeof(T));
    return static_cast<T*>(p);
}
} // namespace

extern "C" int nbref_real_bytes()
{
    return static_cast<int>(sizeof(real));
}

/* The sizes of the list structures as the reference lays them out (tests compare them with include/fepb200_nb.h). */
extern "C" void nbref_struct_sizes(int* out)
{
    out[0] = sizeof(nbnxn_sci_t);
    out[1] = sizeof(nbnxn_cj_packed_t);
    out[2] = sizeof(nbnxn_excl_t);
    out[3] = c_nbnxnGpuClusterSize;
    out[4] = c_nbnxnGpuNumClusterPerSupercluster;
    out[5] = c_nbnxnGpuJgroupSize;
    out[6] = c_nbnxnGpuClusterpairSplit;
}

/* One call of the reference kernel on zeroed outputs.
 *   xq      double[4 natoms] (x, y, z, q -- q already masked for perturbed atoms, atomdata.cpp:930-964)
 *   type    int[natoms]; nbfp double[2 ntype^2] = {6 C6, 12 C12}
 *   sci     the bytes of nbnxn_sci_t[nsci]; cj of nbnxn_cj_packed_t[ncj]; excl of nbnxn_excl_t[nexcl]
 *   f double[3 natoms], fshift double[135], vc / vvdw double[1] */
extern "C" int nbref_run(int natoms, const double* xq, const int* type, int ntype, const double* nbfp,
                         const nbref_params* p, int nsci, const void* sci, int ncj, const void* cj, int nexcl,
                         const void* excl, const double* shiftvec, int computeEnergy, double* f, double* fshift,
                         double* vc, double* vvdw)
{
    nbnxn_atomdata_t* nbat = zeroedStorage<nbnxn_atomdata_t>();
    new (&nbat->params_.nbfp) gmx::HostVector<real>(nbfp, nbfp + 2 * ntype * ntype);
    new (&nbat->params_.type) gmx::HostVector<int>(type, type + natoms);
    nbat->params_.numTypes = ntype;
    new (&nbat->x_) gmx::HostVector<real>(xq, xq + 4 * static_cast<size_t>(natoms));
    nbat->XFormat = nbatXYZQ;
    nbat->FFormat = nbatXYZ;
    nbat->xstride = STRIDE_XYZQ;
    nbat->fstride = STRIDE_XYZ;

    NbnxnPairlistGpu* nbl = zeroedStorage<NbnxnPairlistGpu>();
    nbl->na_ci            = c_nbnxnGpuClusterSize;
    nbl->na_cj            = c_nbnxnGpuClusterSize;
    nbl->na_sc            = c_nbnxnGpuClusterSize * c_nbnxnGpuNumClusterPerSupercluster;
    nbl->rlist            = p->rlist;
    new (&nbl->sci) gmx::HostVector<nbnxn_sci_t>(nsci);
    std::memcpy(nbl->sci.data(), sci, sizeof(nbnxn_sci_t) * nsci);
    new (&nbl->cjPacked.list_) gmx::HostVector<nbnxn_cj_packed_t>(ncj);
    std::memcpy(nbl->cjPacked.list_.data(), cj, sizeof(nbnxn_cj_packed_t) * ncj);
    new (&nbl->excl) gmx::HostVector<nbnxn_excl_t>(nexcl);
    std::memcpy(nbl->excl.data(), excl, sizeof(nbnxn_excl_t) * nexcl);

    interaction_const_t ic;
    ic.eeltype                  = static_cast<CoulombInteractionType>(p->eeltype);
    ic.epsfac                   = p->epsfac;
    ic.rcoulomb                 = p->rcoulomb;
    ic.rvdw                     = p->rvdw;
    ic.reactionFieldCoefficient = p->reactionFieldCoefficient;
    ic.reactionFieldShift       = p->reactionFieldShift;
    ic.sh_ewald                 = p->sh_ewald;
    ic.ewaldcoeff_q             = p->ewaldcoeff_q;
    ic.dispersion_shift.cpot    = p->dispersion_shift_cpot;
    ic.repulsion_shift.cpot     = p->repulsion_shift_cpot;
    ic.coulombEwaldTables       = std::make_unique<EwaldCorrectionTables>();
    ic.coulombEwaldTables->scale = p->tab_scale;
    if (p->tableF != nullptr)
    {
        ic.coulombEwaldTables->tableF.assign(p->tableF, p->tableF + p->tab_size);
    }

    std::vector<gmx::RVec> sv(gmx::c_numShiftVectors);
    for (int s = 0; s < gmx::c_numShiftVectors; s++)
    {
        sv[s] = { static_cast<real>(shiftvec[3 * s]), static_cast<real>(shiftvec[3 * s + 1]),
                  static_cast<real>(shiftvec[3 * s + 2]) };
    }
    gmx::StepWorkload stepWork;
    stepWork.computeEnergy = computeEnergy != 0;
    stepWork.computeVirial = true;

    std::vector<real> fr(3 * static_cast<size_t>(natoms), 0), fsh(3 * gmx::c_numShiftVectors, 0);
    real              vcr = 0, vvr = 0;
    nbnxn_kernel_gpu_ref(nbl, nbat, &ic, sv, stepWork, enbvClearFYes, fr, fsh.data(), &vcr, &vvr);

    for (size_t k = 0; k < fr.size(); k++)
    {
        f[k] = fr[k];
    }
    for (size_t k = 0; k < fsh.size(); k++)
    {
        fshift[k] = fsh[k];
    }
    *vc   = vcr;
    *vvdw = vvr;

    nbl->excl.~vector();
    nbl->cjPacked.list_.~vector();
    nbl->sci.~vector();
    nbat->x_.~vector();
    nbat->params_.type.~vector();
    nbat->params_.nbfp.~vector();
    std::free(nbl);
    std::free(nbat);
    return 0;
}
