/*
 * oracle/ref_build/harness.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" driver around the UNMODIFIED reference CPU kernel
 *   gmx_nb_free_energy_kernel()   /root/reference/src/gromacs/gmxlib/nonbonded/nb_free_energy.cpp:1367
 * which the Makefile next to this file compiles from where it lies under /root/reference
 * (nothing from the reference is copied into this repository).  The result,
 * oracle/_ref/libfepref_{sp,dp}*.so, is used only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * The driver re-creates, in our own words, the control flow of
 *   dispatchFreeEnergyKernel()    src/gromacs/nbnxm/freeenergydispatch.cpp:147-308
 * (one t_nblist per OpenMP thread, split with the rule of balance_fep_lists(),
 * pairlist.cpp:2786-2838; per-thread force/energy buffers that are cleared, filled and then
 * reduced; L+1 energy-only foreign-lambda passes) so that both results and CPU timings
 * correspond to what `mdrun -nb cpu -fep cpu -ntomp T` does on this path.
 */
#include "config.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <memory>
#include <vector>

#include "gromacs/gmxlib/nonbonded/nb_free_energy.h"
#include "gromacs/gmxlib/nonbonded/nonbonded.h"
#include "gromacs/gmxlib/nrnb.h"
#include "gromacs/math/arrayrefwithpadding.h"
#include "gromacs/math/vectypes.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/md_enums.h"
#include "gromacs/mdtypes/nblist.h"
#include "gromacs/utility/arrayref.h"
#include "gromacs/utility/alignedallocator.h"
#include "gromacs/utility/real.h"

#include "fepb200.h"

/* ---- the two symbols nb_free_energy.o needs from the rest of libgromacs ---------------- */
void atomicNrnbIncrement(t_nrnb* nrnb, int index, int increment)
{
#pragma omp atomic
    nrnb->n[index] += increment;
}
namespace gmx
{
// interaction_const_t's (never filled) Ewald table vectors reference this deallocator
void AlignedAllocationPolicy::free(void* p)
{
    std::free(p);
}
namespace internal
{
[[noreturn]] void assertHandler(const char* condition, const char* msg, const char* func,
                                const std::filesystem::path& file, int line)
{
    std::fprintf(stderr, "reference assertion failed: %s (%s) in %s at %s:%d\n", condition, msg,
                 func, file.c_str(), line);
    std::abort();
}
} // namespace internal
} // namespace gmx

/* Same fields as fepb200_params (include/fepb200.h) but in double, so that the
 * double-precision build can be fed the exact inputs of the reference's unit test. */
extern "C" struct fepref_params
{
    int    eeltype, vdwtype, vdw_modifier;
    double epsfac, rcoulomb, rvdw, rvdw_switch, reactionFieldCoefficient, reactionFieldShift;
    double sh_ewald, sh_lj_ewald, ewaldcoeff_q, ewaldcoeff_lj, dispersion_shift_cpot, repulsion_shift_cpot;
    int    softcoreType;
    double alphaVdw, alphaCoulomb;
    int    lambdaPower;
    double sigma6WithInvalidSigma, sigma6Minimum, gapsysScaleLinpointVdW, gapsysScaleLinpointCoul,
            gapsysSigma6VdW;
};

namespace
{

constexpr int c_numShift  = FEPB200_NUM_SHIFT_VECTORS;
constexpr int c_blockSize = 32; // atoms per reduction block, as in ThreadedForceBuffer

void fillInteractionConst(const fepref_params& p, interaction_const_t* ic)
{
    ic->eeltype                  = static_cast<CoulombInteractionType>(p.eeltype);
    ic->vdwtype                  = static_cast<VanDerWaalsType>(p.vdwtype);
    ic->vdw_modifier             = static_cast<InteractionModifiers>(p.vdw_modifier);
    ic->epsfac                   = p.epsfac;
    ic->rcoulomb                 = p.rcoulomb;
    ic->rvdw                     = p.rvdw;
    ic->rvdw_switch              = p.rvdw_switch;
    ic->reactionFieldCoefficient = p.reactionFieldCoefficient;
    ic->reactionFieldShift       = p.reactionFieldShift;
    ic->sh_ewald                 = p.sh_ewald;
    ic->sh_lj_ewald              = p.sh_lj_ewald;
    ic->ewaldcoeff_q             = p.ewaldcoeff_q;
    ic->ewaldcoeff_lj            = p.ewaldcoeff_lj;
    ic->dispersion_shift.cpot    = p.dispersion_shift_cpot;
    ic->repulsion_shift.cpot     = p.repulsion_shift_cpot;
    // SoftCoreParameters only has a constructor from t_lambda that lives in another
    // translation unit; it is a trivially destructible aggregate of scalars, so we
    // allocate raw storage and assign the fields directly.
    using SC = interaction_const_t::SoftCoreParameters;
    SC* sc   = static_cast<SC*>(::operator new(sizeof(SC)));
    sc->alphaVdw                = p.alphaVdw;
    sc->alphaCoulomb            = p.alphaCoulomb;
    sc->lambdaPower             = p.lambdaPower;
    sc->sigma6WithInvalidSigma  = p.sigma6WithInvalidSigma;
    sc->sigma6Minimum           = p.sigma6Minimum;
    sc->softcoreType            = static_cast<SoftcoreType>(p.softcoreType);
    sc->gapsysScaleLinpointVdW  = p.gapsysScaleLinpointVdW;
    sc->gapsysScaleLinpointCoul = p.gapsysScaleLinpointCoul;
    sc->gapsysSigma6VdW         = p.gapsysSigma6VdW;
    ic->softCoreParameters.reset(sc);
}

struct ThreadBuffers
{
    std::vector<gmx::RVec> f;       // natoms + padding
    std::vector<gmx::RVec> fshift;  // 45
    std::vector<real>      vc, vv;  // G
    std::vector<real>      dvdl;    // 7
    std::vector<int>       blocks;  // reduction blocks this thread touches
};

template<typename T>
std::vector<real> toReal(const T* src, size_t n)
{
    std::vector<real> v(n);
    for (size_t i = 0; i < n; i++)
    {
        v[i] = static_cast<real>(src[i]);
    }
    return v;
}

double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" int fepref_real_bytes()
{
    return static_cast<int>(sizeof(real));
}

extern "C" const char* fepref_simd_string()
{
    return GMX_SIMD_STRING;
}

/* Runs `repeats` force(+energy) passes at the current lambda, and, when
 * FEPB200_DO_FOREIGNLAMBDA is set, `repeats` sweeps of the L+1 energy-only passes.
 * Outputs are overwritten with the result of ONE step.  seconds[0] / seconds[1] receive the
 * best wall time of a current-lambda pass / a foreign sweep, including clear and reduction. */
extern "C" int fepref_dispatch(const fepref_params* p, int useSimd, int nthreads, int ntype,
                               const double* nbfp_in, const double* nbfpGrid_in, int natoms,
                               const double* x_in, const double* qA_in, const double* qB_in,
                               const int* typeA_in, const int* typeB_in, const double* shiftvec_in,
                               int nri, const int* iinr, const int* gid, const int* shift,
                               const int* jindex, const int* jjnr, const int* exclFep, int numGroupPairs,
                               int flags, const double* lambda_in, int nForeign,
                               const double* allLambdaCoul, const double* allLambdaVdw, double* fOut,
                               double* fshiftOut, double* vcOut, double* vvOut, double* dvdlOut,
                               double* foreignE, double* foreignDvdl, int repeats, double* seconds)
{
    if (nthreads < 1)
    {
        nthreads = 1;
    }
    interaction_const_t ic;
    fillInteractionConst(*p, &ic);

    const std::vector<real> nbfp = toReal(nbfp_in, 2 * size_t(ntype) * ntype);
    std::vector<real>       nbfpGrid(2 * size_t(ntype) * ntype, 0);
    if (nbfpGrid_in)
    {
        nbfpGrid = toReal(nbfpGrid_in, 2 * size_t(ntype) * ntype);
    }
    const std::vector<real> qA = toReal(qA_in, natoms);
    const std::vector<real> qB = toReal(qB_in, natoms);
    const std::vector<int>  typeA(typeA_in, typeA_in + natoms);
    const std::vector<int>  typeB(typeB_in, typeB_in + natoms);

    const int              npad = natoms + 16; // SIMD gathers may read one element past the end
    std::vector<gmx::RVec> x(npad, gmx::RVec(0, 0, 0));
    for (int a = 0; a < natoms; a++)
    {
        x[a] = gmx::RVec(x_in[3 * a], x_in[3 * a + 1], x_in[3 * a + 2]);
    }
    std::vector<gmx::RVec> shiftvec(c_numShift);
    for (int s = 0; s < c_numShift; s++)
    {
        shiftvec[s] = gmx::RVec(shiftvec_in[3 * s], shiftvec_in[3 * s + 1], shiftvec_in[3 * s + 2]);
    }

    /* Split the list over threads: walk the entries once, move on to the next thread when
     * adding the entry would overshoot the per-thread pair target by more than the
     * current shortfall (the balance_fep_lists rule). */
    std::vector<t_nblist> lists(nthreads);
    {
        const long nrjTot    = jindex[nri];
        const long nrjTarget = (nrjTot + nthreads - 1) / nthreads;
        int        dest      = 0;
        for (auto& l : lists)
        {
            l.jindex.push_back(0);
        }
        for (int n = 0; n < nri; n++)
        {
            const int nrj = jindex[n + 1] - jindex[n];
            t_nblist* l   = &lists[dest];
            if (dest + 1 < nthreads && l->nrj > 0 && l->nrj + nrj - nrjTarget > nrjTarget - l->nrj)
            {
                dest++;
                l = &lists[dest];
            }
            l->iinr.push_back(iinr[n]);
            l->gid.push_back(gid[n]);
            l->shift.push_back(shift[n]);
            for (int k = jindex[n]; k < jindex[n + 1]; k++)
            {
                l->jjnr.push_back(jjnr[k]);
                l->excl_fep.push_back(exclFep ? exclFep[k] : 1);
            }
            l->nrj += nrj;
            l->nri++;
            l->jindex.push_back(l->nrj);
        }
        for (auto& l : lists)
        {
            // the SIMD kernel loads jjnr[k] for the padded lanes of the last chunk
            l.jjnr.resize(l.jjnr.size() + 16, l.jjnr.empty() ? 0 : l.jjnr.back());
            l.maxnri = l.nri;
            l.maxnrj = l.nrj;
        }
    }

    const int                  numBlocks = (natoms + c_blockSize - 1) / c_blockSize;
    std::vector<ThreadBuffers> tb(nthreads);
    for (int th = 0; th < nthreads; th++)
    {
        tb[th].f.assign(npad, gmx::RVec(0, 0, 0));
        tb[th].fshift.assign(c_numShift, gmx::RVec(0, 0, 0));
        tb[th].vc.assign(numGroupPairs, 0);
        tb[th].vv.assign(numGroupPairs, 0);
        tb[th].dvdl.assign(FEPB200_NUM_LAMBDA_COMPONENTS, 0);
        std::vector<char> used(numBlocks, 0);
        const t_nblist&   l = lists[th];
        for (int n = 0; n < l.nri; n++)
        {
            used[l.iinr[n] / c_blockSize] = 1;
            for (int k = l.jindex[n]; k < l.jindex[n + 1]; k++)
            {
                used[l.jjnr[k] / c_blockSize] = 1;
            }
        }
        for (int b = 0; b < numBlocks; b++)
        {
            if (used[b])
            {
                tb[th].blocks.push_back(b);
            }
        }
    }
    // blocks touched by any thread, for the output reduction
    std::vector<int> allBlocks;
    {
        std::vector<char> used(numBlocks, 0);
        for (auto& t : tb)
        {
            for (int b : t.blocks)
            {
                used[b] = 1;
            }
        }
        for (int b = 0; b < numBlocks; b++)
        {
            if (used[b])
            {
                allBlocks.push_back(b);
            }
        }
    }

    std::vector<real> lambda = toReal(lambda_in, FEPB200_NUM_LAMBDA_COMPONENTS);
    t_nrnb            nrnb;

    const gmx::ArrayRefWithPadding<const gmx::RVec> coords(x.data(), x.data() + natoms,
                                                           x.data() + npad);

    auto clearThread = [&](ThreadBuffers& t) {
        for (int b : t.blocks)
        {
            const int a1 = std::min(natoms, (b + 1) * c_blockSize);
            for (int a = b * c_blockSize; a < a1; a++)
            {
                t.f[a] = gmx::RVec(0, 0, 0);
            }
        }
        std::fill(t.fshift.begin(), t.fshift.end(), gmx::RVec(0, 0, 0));
        std::fill(t.vc.begin(), t.vc.end(), real(0));
        std::fill(t.vv.begin(), t.vv.end(), real(0));
        std::fill(t.dvdl.begin(), t.dvdl.end(), real(0));
    };

    std::vector<gmx::RVec> fSum(natoms, gmx::RVec(0, 0, 0));
    std::vector<gmx::RVec> fshiftSum(c_numShift);
    std::vector<real>      vcSum(numGroupPairs), vvSum(numGroupPairs);
    real                   dvdlSum[2];

    double bestForce = 1e30, bestForeign = 1e30;
    repeats          = std::max(repeats, 1);

    for (int rep = 0; rep < repeats; rep++)
    {
        const double t0 = now();
#pragma omp parallel for schedule(static) num_threads(nthreads)
        for (int th = 0; th < nthreads; th++)
        {
            ThreadBuffers& t = tb[th];
            clearThread(t);
            gmx::ArrayRefWithPadding<gmx::RVec> fRef(t.f.data(), t.f.data() + natoms,
                                                     t.f.data() + npad);
            gmx_nb_free_energy_kernel(lists[th], coords, useSimd != 0, ntype, ic, shiftvec, nbfp,
                                      nbfpGrid, qA, qB, typeA, typeB, flags & ~FEPB200_DO_FOREIGNLAMBDA,
                                      lambda, &nrnb, fRef,
                                      reinterpret_cast<rvec*>(t.fshift.data()), t.vc, t.vv, t.dvdl);
        }
        // reduction over threads (serial over threads, parallel over atom blocks)
        if (flags & FEPB200_DO_FORCE)
        {
            const int nb = static_cast<int>(allBlocks.size());
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int ib = 0; ib < nb; ib++)
            {
                const int b  = allBlocks[ib];
                const int a1 = std::min(natoms, (b + 1) * c_blockSize);
                for (int a = b * c_blockSize; a < a1; a++)
                {
                    gmx::RVec s(0, 0, 0);
                    for (int th = 0; th < nthreads; th++)
                    {
                        s += tb[th].f[a];
                    }
                    fSum[a] = s;
                }
            }
        }
        for (int s = 0; s < c_numShift; s++)
        {
            gmx::RVec v(0, 0, 0);
            for (int th = 0; th < nthreads; th++)
            {
                v += tb[th].fshift[s];
            }
            fshiftSum[s] = v;
        }
        for (int g = 0; g < numGroupPairs; g++)
        {
            real a = 0, b = 0;
            for (int th = 0; th < nthreads; th++)
            {
                a += tb[th].vc[g];
                b += tb[th].vv[g];
            }
            vcSum[g] = a;
            vvSum[g] = b;
        }
        dvdlSum[0] = dvdlSum[1] = 0;
        for (int th = 0; th < nthreads; th++)
        {
            dvdlSum[0] += tb[th].dvdl[FEPB200_LAMBDA_COUL];
            dvdlSum[1] += tb[th].dvdl[FEPB200_LAMBDA_VDW];
        }
        bestForce = std::min(bestForce, now() - t0);
    }
    for (int a = 0; a < natoms; a++)
    {
        for (int d = 0; d < 3; d++)
        {
            fOut[3 * a + d] = fSum[a][d];
        }
    }
    for (int s = 0; s < c_numShift; s++)
    {
        for (int d = 0; d < 3; d++)
        {
            fshiftOut[3 * s + d] = fshiftSum[s][d];
        }
    }
    for (int g = 0; g < numGroupPairs; g++)
    {
        vcOut[g] = vcSum[g];
        vvOut[g] = vvSum[g];
    }
    dvdlOut[0] = dvdlSum[0];
    dvdlOut[1] = dvdlSum[1];

    if ((flags & FEPB200_DO_FOREIGNLAMBDA) && foreignE != nullptr)
    {
        const int kernelFlags = (flags & ~(FEPB200_DO_FORCE | FEPB200_DO_SHIFTFORCE))
                                | FEPB200_DO_FOREIGNLAMBDA | FEPB200_DO_POTENTIAL;
        for (int rep = 0; rep < repeats; rep++)
        {
            const double t0 = now();
            for (int i = 0; i < 1 + nForeign; i++)
            {
                std::vector<real> lam = lambda;
                if (i > 0)
                {
                    // only the Coul and Vdw components are read by the kernel
                    lam[FEPB200_LAMBDA_COUL] = allLambdaCoul[i - 1];
                    lam[FEPB200_LAMBDA_VDW]  = allLambdaVdw[i - 1];
                }
#pragma omp parallel for schedule(static) num_threads(nthreads)
                for (int th = 0; th < nthreads; th++)
                {
                    ThreadBuffers& t = tb[th];
                    std::fill(t.vc.begin(), t.vc.end(), real(0));
                    std::fill(t.vv.begin(), t.vv.end(), real(0));
                    std::fill(t.dvdl.begin(), t.dvdl.end(), real(0));
                    gmx_nb_free_energy_kernel(lists[th], coords, useSimd != 0, ntype, ic, shiftvec,
                                              nbfp, nbfpGrid, qA, qB, typeA, typeB, kernelFlags, lam,
                                              &nrnb, gmx::ArrayRefWithPadding<gmx::RVec>(), nullptr,
                                              t.vc, t.vv, t.dvdl);
                }
                // reduce group-pair energies over threads in `real`, then sum_epot
                real epot = 0, vcTot = 0, vvTot = 0, dc = 0, dv = 0;
                for (int g = 0; g < numGroupPairs; g++)
                {
                    real a = 0, b = 0;
                    for (int th = 0; th < nthreads; th++)
                    {
                        a += tb[th].vc[g];
                        b += tb[th].vv[g];
                    }
                    vcTot += a;
                    vvTot += b;
                }
                epot = vcTot + vvTot;
                for (int th = 0; th < nthreads; th++)
                {
                    dc += tb[th].dvdl[FEPB200_LAMBDA_COUL];
                    dv += tb[th].dvdl[FEPB200_LAMBDA_VDW];
                }
                foreignE[i]            = epot;
                foreignDvdl[2 * i]     = dc;
                foreignDvdl[2 * i + 1] = dv;
            }
            bestForeign = std::min(bestForeign, now() - t0);
        }
    }
    else
    {
        bestForeign = 0;
    }
    if (seconds)
    {
        seconds[0] = bestForce;
        seconds[1] = bestForeign;
    }
    return 0;
}
