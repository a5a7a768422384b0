/*
 * oracle/ref_build/fork_cuda/harness.cu -- TEST / MEASUREMENT INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" driver around the UNMODIFIED CUDA FEP kernels of the reference fork,
 *   nbnxn_fep_kernel_*_{F,VF}_cuda        /root/reference/src/gromacs/nbnxm/cuda/nbnxm_fep_cuda_kernel.cuh
 *   nbnxn_foreign_fep_kernel_*_V_cuda     /root/reference/src/gromacs/nbnxm/cuda/nbnxm_foreign_fep_cuda_kernel.cuh
 * which the Makefile next to this file compiles for sm_100a from where they lie under
 * /root/reference (#include below; nothing from the reference is copied into this repository).
 * The result, oracle/_ref/libfepfork_cuda.so, lets tests/ and bench.py put the fork's own GPU
 * kernels beside ours on the same B200 and the same problem: "the kernel to beat" of SURVEY 8a-9.
 *
 * The driver re-creates, in our own words, what the fork's host glue does around those kernels:
 *   gpu_init_atomdata / cuda_copy_fepparams / gpu_init_feppairlist
 *                               src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:491-536,761-871,990-1040
 *   the launches in gpu_launch_kernel  src/gromacs/nbnxm/cuda/nbnxm_cuda.cu:755-851
 *     (64 threads per block, one warp per i-entry; foreign kernel when n_lambda > 0 and soft-core)
 * Flavours instantiated: analytical Ewald or reaction-field electrostatics with plain cut-off LJ
 * (potential shift), i.e. what BASELINE.json's Beutler configurations select.  The fork has no
 * Gapsys soft-core, no energy groups and no separate outputs per energy-group pair.
 *
 * Same argument list as fepref_dispatch() (oracle/ref_build/harness.cpp) / fep_oracle_dispatch():
 * doubles in, doubles out (the kernels compute in float).  seconds[0] = device time of the
 * current-lambda kernel, seconds[1] = of the foreign-lambda kernel (CUDA events, best of `repeats`).
 */
#include "gromacs/gpu_utils/cudautils.cuh"
#include "gromacs/gpu_utils/typecasts.cuh"

#include "nbnxm_cuda_kernel_utils.cuh"
#include "nbnxm_cuda_types.h"

/* ---- the fork's kernels, generated the way nbnxm_fep_cuda_kernels.cuh does it ------------------ */
#define EL_EWALD_ANA
#define NB_FEP_KERNEL_FUNC_NAME(x, ...) x##_ElecEw_VdwLJ##__VA_ARGS__
#define NB_FOREIGN_FEP_KERNEL_FUNC_NAME(x, ...) x##_ElecEw_VdwLJ##__VA_ARGS__
#include "nbnxm_fep_cuda_kernel.cuh" /* F */
#define CALC_ENERGIES
#include "nbnxm_fep_cuda_kernel.cuh" /* VF */
#undef CALC_ENERGIES
#include "nbnxm_foreign_fep_cuda_kernel.cuh" /* foreign V */
#undef NB_FEP_KERNEL_FUNC_NAME
#undef NB_FOREIGN_FEP_KERNEL_FUNC_NAME
#undef EL_EWALD_ANA

#define EL_RF
#define NB_FEP_KERNEL_FUNC_NAME(x, ...) x##_ElecRF_VdwLJ##__VA_ARGS__
#define NB_FOREIGN_FEP_KERNEL_FUNC_NAME(x, ...) x##_ElecRF_VdwLJ##__VA_ARGS__
#include "nbnxm_fep_cuda_kernel.cuh"
#define CALC_ENERGIES
#include "nbnxm_fep_cuda_kernel.cuh"
#undef CALC_ENERGIES
#include "nbnxm_foreign_fep_cuda_kernel.cuh"
#undef NB_FEP_KERNEL_FUNC_NAME
#undef NB_FOREIGN_FEP_KERNEL_FUNC_NAME
#undef EL_RF

#include <cstdio>
#include <cstring>
#include <vector>

/* same fields as fepref_params in ../harness.cpp */
extern "C" struct fepfork_params
{
    int    eeltype, vdwtype, vdw_modifier;
    double epsfac, rcoulomb, rvdw, rvdw_switch, krf, crf;
    double sh_ewald, sh_lj_ewald, ewaldcoeff_q, ewaldcoeff_lj, dispersion_cpot, repulsion_cpot;
    int    softcoreType;
    double alphaVdw, alphaCoulomb;
    int    lambdaPower;
    double sigma6WithInvalidSigma, sigma6Minimum, gapsysScaleVdW, gapsysScaleCoul, gapsysSigma6VdW;
};
typedef fepfork_params Params;

namespace
{

constexpr int DO_FORCE = 1 << 1, DO_SHIFTFORCE = 1 << 2, DO_FOREIGNLAMBDA = 1 << 3, DO_POTENTIAL = 1 << 4;
constexpr int NUM_SHIFT = 45;

#define CK(call)                                                                                       \
    do                                                                                                 \
    {                                                                                                  \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
        {                                                                                              \
            std::fprintf(stderr, "fepfork: %s failed: %s\n", #call, cudaGetErrorString(e_));           \
            return -2;                                                                                 \
        }                                                                                              \
    } while (0)

template<typename T>
struct Dev
{
    T* p = nullptr;
    ~Dev() { cudaFree(p); }
    cudaError_t upload(const std::vector<T>& h, size_t pad = 0)
    {
        cudaError_t e = cudaMalloc(&p, sizeof(T) * (h.size() + pad + 1));
        if (e != cudaSuccess)
        {
            return e;
        }
        e = cudaMemset(p, 0, sizeof(T) * (h.size() + pad + 1));
        if (e != cudaSuccess || h.empty())
        {
            return e;
        }
        return cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
    }
    cudaError_t zeros(size_t n)
    {
        cudaError_t e = cudaMalloc(&p, sizeof(T) * (n + 1));
        return e != cudaSuccess ? e : cudaMemset(p, 0, sizeof(T) * (n + 1));
    }
};

} // namespace

extern "C" const char* fepfork_describe()
{
    return "reference fork CUDA FEP kernels (nbnxm_fep_cuda_kernel.cuh, nbnxm_foreign_fep_cuda_kernel.cuh) compiled in place for sm_100a";
}

/* 0 = supported, else why not (the fork's kernels cover less than the CPU kernel) */
extern "C" int fepfork_supports(const Params* p, int ngrp)
{
    const bool ewald = p->eeltype == 3 || p->eeltype == 4; /* PME, Ewald */
    const bool rf    = p->eeltype == 0 || p->eeltype == 1 || p->eeltype == 16;
    if (!(ewald || rf))
    {
        return 1;
    }
    if (p->softcoreType != 0) /* Gapsys */
    {
        return 2;
    }
    if (p->vdwtype != 0 || p->vdw_modifier == 3 || p->vdw_modifier == 5) /* LJ-PME, pot-switch, force-switch */
    {
        return 3;
    }
    if (ngrp != 1)
    {
        return 4;
    }
    if (p->rcoulomb != p->rvdw) /* would need the VDW_CUTOFF_CHECK flavours */
    {
        return 5;
    }
    return 0;
}

extern "C" int fepfork_dispatch(const Params* p, int /*use_simd*/, int /*nthreads*/, int ntype, const double* nbfp,
                                const double* /*nbfp_grid*/, int natoms, const double* x, const double* qA,
                                const double* qB, const int* typeA, const int* typeB, const double* shiftvec, int nri,
                                const int* iinr, const int* /*gid*/, const int* shift, const int* jindex, const int* jjnr,
                                const int* excl, int ngrp, int flags, const double* lambda, int nforeign,
                                const double* all_lambda_coul, const double* all_lambda_vdw, double* f, double* fshift,
                                double* Vc, double* Vv, double* dvdl, double* foreign_e, double* foreign_dvdl, int repeats,
                                double* seconds)
{
    if (fepfork_supports(p, ngrp) != 0)
    {
        return -4;
    }
    if (nri <= 0)
    {
        return 0; /* the fork does not launch on an empty list (nbnxm_cuda.cu:758-766) */
    }
    const bool ewald = p->eeltype == 3 || p->eeltype == 4;
    const int  nrj   = jindex[nri];

    /* atom data: gpu_init_atomdata FEP part (xq, q4 = {qA, qB}, atomTypes4 = {typeA, typeB}) */
    std::vector<float4> h_xq(natoms), h_q4(natoms);
    std::vector<int4>   h_t4(natoms);
    for (int a = 0; a < natoms; a++)
    {
        h_xq[a] = make_float4((float)x[3 * a], (float)x[3 * a + 1], (float)x[3 * a + 2], (float)qA[a]);
        h_q4[a] = make_float4((float)qA[a], (float)qB[a], 0.0F, 0.0F);
        h_t4[a] = make_int4(typeA[a], typeB[a], 0, 0);
    }
    std::vector<float2> h_nbfp((size_t)ntype * ntype);
    for (size_t i = 0; i < h_nbfp.size(); i++)
    {
        h_nbfp[i] = make_float2((float)nbfp[2 * i], (float)nbfp[2 * i + 1]); /* 6*C6, 12*C12 */
    }
    std::vector<float> h_shift(3 * NUM_SHIFT), h_alc(nforeign), h_alv(nforeign);
    for (int i = 0; i < 3 * NUM_SHIFT; i++)
    {
        h_shift[i] = (float)shiftvec[i];
    }
    for (int i = 0; i < nforeign; i++)
    {
        h_alc[i] = (float)all_lambda_coul[i];
        h_alv[i] = (float)all_lambda_vdw[i];
    }
    std::vector<int> h_iinr(iinr, iinr + nri), h_shiftidx(shift, shift + nri), h_jindex(jindex, jindex + nri + 1),
            h_jjnr(jjnr, jjnr + nrj), h_excl(nrj, 1);
    if (excl)
    {
        h_excl.assign(excl, excl + nrj);
    }

    Dev<float4> d_xq, d_q4;
    Dev<int4>   d_t4;
    Dev<float2> d_nbfp;
    Dev<float>  d_f, d_fshift, d_shiftvec, d_e, d_alc, d_alv;
    Dev<int>    d_iinr, d_shiftidx, d_jindex, d_jjnr, d_excl;
    CK(d_xq.upload(h_xq));
    CK(d_q4.upload(h_q4));
    CK(d_t4.upload(h_t4));
    CK(d_nbfp.upload(h_nbfp));
    CK(d_shiftvec.upload(h_shift));
    CK(d_alc.upload(h_alc));
    CK(d_alv.upload(h_alv));
    CK(d_iinr.upload(h_iinr));
    CK(d_shiftidx.upload(h_shiftidx));
    CK(d_jindex.upload(h_jindex));
    /* the foreign kernel reads jjnr / excl_fep of all 32 lanes of the last trip of an entry
     * (nbnxm_foreign_fep_cuda_kernel.cuh:262-264): pad, so that those reads stay inside the buffers */
    CK(d_jjnr.upload(h_jjnr, 32));
    CK(d_excl.upload(h_excl, 32));
    CK(d_f.zeros(3 * (size_t)natoms));
    CK(d_fshift.zeros(3 * NUM_SHIFT));
    const int L = nforeign;
    /* eLJ eElec dvdlLJ dvdlElec | eLJForeign[L+1] eElecForeign[L+1] dvdlLJForeign[L+1] dvdlElecForeign[L+1] */
    const size_t n_e = 4 + 4 * (size_t)(L + 1);
    CK(d_e.zeros(n_e));

    /* the LJ parameter table is read through a texture object on this architecture
     * (cuda_arch_utils.cuh:76-81; initParamLookupTable, gpu_utils/cudautils.cu) */
    cudaTextureObject_t nbfp_tex = 0;
    {
        cudaResourceDesc rd;
        std::memset(&rd, 0, sizeof(rd));
        rd.resType                = cudaResourceTypeLinear;
        rd.res.linear.devPtr      = d_nbfp.p;
        rd.res.linear.desc        = cudaCreateChannelDesc<float2>();
        rd.res.linear.sizeInBytes = h_nbfp.size() * sizeof(float2);
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&nbfp_tex, &rd, &td, nullptr));
    }

    NBAtomDataGpu adat{};
    adat.numAtoms = adat.numAtomsLocal = adat.numAtomsAlloc = natoms;
    adat.xq                                                 = d_xq.p;
    adat.q4                                                 = d_q4.p;
    adat.f                                                  = reinterpret_cast<Float3*>(d_f.p);
    adat.fShift                                             = reinterpret_cast<Float3*>(d_fshift.p);
    adat.shiftVec                                           = reinterpret_cast<Float3*>(d_shiftvec.p);
    adat.shiftVecUploaded                                   = true;
    adat.eLJ                                                = d_e.p + 0;
    adat.eElec                                              = d_e.p + 1;
    adat.dvdlLJ                                             = d_e.p + 2;
    adat.dvdlElec                                           = d_e.p + 3;
    adat.eLJForeign                                         = d_e.p + 4;
    adat.eElecForeign                                       = d_e.p + 4 + (L + 1);
    adat.dvdlLJForeign                                      = d_e.p + 4 + 2 * (L + 1);
    adat.dvdlElecForeign                                    = d_e.p + 4 + 3 * (L + 1);
    adat.numTypes                                           = ntype;
    adat.atomTypes4                                         = d_t4.p;

    NBParamGpu nbp{};
    nbp.elecType = ewald ? Nbnxm::ElecType::EwaldAna : Nbnxm::ElecType::RF;
    nbp.vdwType  = Nbnxm::VdwType::Cut;
    /* set_cutoff_parameters(), nbnxm_gpu_data_mgmt.cpp:201-223 */
    nbp.ewald_beta            = (float)p->ewaldcoeff_q;
    nbp.sh_ewald              = (float)p->sh_ewald;
    nbp.epsfac                = (float)p->epsfac;
    nbp.two_k_rf              = (float)(2.0 * p->krf);
    nbp.c_rf                  = (float)p->crf;
    nbp.rvdw_sq               = (float)(p->rvdw * p->rvdw);
    nbp.rcoulomb_sq           = (float)(p->rcoulomb * p->rcoulomb);
    nbp.rlistOuter_sq         = nbp.rcoulomb_sq;
    nbp.rlistInner_sq         = nbp.rcoulomb_sq;
    nbp.useDynamicPruning     = false;
    nbp.sh_lj_ewald           = (float)p->sh_lj_ewald;
    nbp.ewaldcoeff_lj         = (float)p->ewaldcoeff_lj;
    nbp.rvdw_switch           = (float)p->rvdw_switch;
    nbp.dispersion_shift.cpot = (float)p->dispersion_cpot;
    nbp.repulsion_shift.cpot  = (float)p->repulsion_cpot;
    nbp.nbfp                  = reinterpret_cast<Float2*>(d_nbfp.p);
    nbp.nbfp_texobj           = nbfp_tex;
    /* cuda_copy_fepparams(), nbnxm_gpu_data_mgmt.cpp:491-536 */
    nbp.bFEP          = true;
    nbp.alpha_coul    = (float)p->alphaCoulomb;
    nbp.alpha_vdw     = (float)p->alphaVdw;
    nbp.lam_power     = p->lambdaPower;
    nbp.sc_sigma6     = (float)p->sigma6WithInvalidSigma;
    nbp.sc_sigma6_min = (float)p->sigma6Minimum;
    nbp.lambda_q      = (float)lambda[2];
    nbp.lambda_v      = (float)lambda[3];
    nbp.allLambdaCoul = d_alc.p;
    nbp.allLambdaVdw  = d_alv.p;

    Nbnxm::gpu_feplist fl{};
    fl.nri = fl.maxnri = nri;
    fl.nrj = fl.maxnrj = nrj;
    fl.iinr            = d_iinr.p;
    fl.shift           = d_shiftidx.p;
    fl.jindex          = d_jindex.p;
    fl.jjnr            = d_jjnr.p;
    fl.excl_fep        = d_excl.p;

    /* launch configuration of nbnxm_cuda.cu:774-787, :806-818 */
    const dim3 block(64, 1, 1);
    const int  nri_per_block = 64 / warp_size;
    const dim3 grid((nri + nri_per_block - 1) / nri_per_block, 1, 1);
    const bool energy  = (flags & DO_POTENTIAL) != 0;
    const bool virial  = (flags & DO_SHIFTFORCE) != 0;
    const bool foreign = L > 0 && (flags & DO_FOREIGNLAMBDA) != 0 && (nbp.alpha_coul != 0 || nbp.alpha_vdw != 0);
    const size_t shmem = (size_t)(L + 1) * 2 * sizeof(float);

    cudaEvent_t ev[3];
    for (auto& e : ev)
    {
        CK(cudaEventCreate(&e));
    }
    float best[2] = { 1e30F, 1e30F };
    if (repeats < 1)
    {
        repeats = 1;
    }
    for (int rep = 0; rep < repeats; rep++)
    {
        CK(cudaMemset(d_f.p, 0, sizeof(float) * 3 * (size_t)natoms));
        CK(cudaMemset(d_fshift.p, 0, sizeof(float) * 3 * NUM_SHIFT));
        CK(cudaMemset(d_e.p, 0, sizeof(float) * n_e));
        CK(cudaEventRecord(ev[0]));
        if (ewald)
        {
            if (energy)
            {
                nbnxn_fep_kernel_ElecEw_VdwLJ_VF_cuda<<<grid, block>>>(adat, nbp, fl, virial);
            }
            else
            {
                nbnxn_fep_kernel_ElecEw_VdwLJ_F_cuda<<<grid, block>>>(adat, nbp, fl, virial);
            }
        }
        else
        {
            if (energy)
            {
                nbnxn_fep_kernel_ElecRF_VdwLJ_VF_cuda<<<grid, block>>>(adat, nbp, fl, virial);
            }
            else
            {
                nbnxn_fep_kernel_ElecRF_VdwLJ_F_cuda<<<grid, block>>>(adat, nbp, fl, virial);
            }
        }
        CK(cudaEventRecord(ev[1]));
        if (foreign)
        {
            if (ewald)
            {
                nbnxn_foreign_fep_kernel_ElecEw_VdwLJ_V_cuda<<<grid, block, shmem>>>(adat, nbp, fl, L);
            }
            else
            {
                nbnxn_foreign_fep_kernel_ElecRF_VdwLJ_V_cuda<<<grid, block, shmem>>>(adat, nbp, fl, L);
            }
        }
        CK(cudaEventRecord(ev[2]));
        CK(cudaEventSynchronize(ev[2]));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
        best[0] = ms < best[0] ? ms : best[0];
        CK(cudaEventElapsedTime(&ms, ev[1], ev[2]));
        best[1] = ms < best[1] ? ms : best[1];
    }
    for (auto& e : ev)
    {
        cudaEventDestroy(e);
    }
    cudaDestroyTextureObject(nbfp_tex);

    /* copy-back, gpu_launch_cpyback (nbnxm_gpu_data_mgmt.cpp:1117 ff.) */
    std::vector<float> h_f(3 * (size_t)natoms), h_fs(3 * NUM_SHIFT), h_e(n_e);
    CK(cudaMemcpy(h_f.data(), d_f.p, sizeof(float) * h_f.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_fs.data(), d_fshift.p, sizeof(float) * h_fs.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_e.data(), d_e.p, sizeof(float) * n_e, cudaMemcpyDeviceToHost));
    if (flags & DO_FORCE)
    {
        for (size_t i = 0; i < h_f.size(); i++)
        {
            f[i] += h_f[i];
        }
        if (virial)
        {
            for (int i = 0; i < 3 * NUM_SHIFT; i++)
            {
                fshift[i] += h_fs[i];
            }
        }
    }
    if (energy)
    {
        Vv[0] += h_e[0];
        Vc[0] += h_e[1];
        dvdl[1] += h_e[2];
        dvdl[0] += h_e[3];
    }
    if (foreign)
    {
        for (int i = 0; i <= L; i++)
        {
            foreign_e[i] += (double)h_e[4 + i] + (double)h_e[4 + (L + 1) + i];
            foreign_dvdl[2 * i + 1] += h_e[4 + 2 * (L + 1) + i];
            foreign_dvdl[2 * i] += h_e[4 + 3 * (L + 1) + i];
        }
    }
    if (seconds)
    {
        seconds[0] = 1e-3 * best[0];
        seconds[1] = foreign ? 1e-3 * best[1] : 0.0;
    }
    return 0;
}
