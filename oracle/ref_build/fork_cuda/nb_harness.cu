/*
 * oracle/ref_build/fork_cuda/nb_harness.cu -- TEST / MEASUREMENT INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" driver around the UNMODIFIED CUDA cluster-pair kernels of the reference (fork),
 *   nbnxn_kernel_ElecEw_VdwLJ_{F,VF}_cuda, nbnxn_kernel_ElecRF_VdwLJ_{F,VF}_cuda
 *                                   /root/reference/src/gromacs/nbnxm/cuda/nbnxm_cuda_kernel.cuh
 * compiled for sm_100a from where they lie under /root/reference (#include below; nothing is copied into this
 * repository).  The result, oracle/_ref/libnbfork_cuda.so, puts the reference's own GPU kernel for the NON-perturbed
 * pairs beside ours (csrc/nb/fep_nb.cu) on the same B200 and the same list: the GPU baseline of SURVEY 8f-3.
 * Launched as gpu_launch_kernel does (nbnxm/cuda/nbnxm_cuda.cu:642-754): one block of 8 x 8 threads per list entry,
 * shared memory per calc_shmem_required_nonbonded (:594-623), no pruning, LJ parameters through the texture object.
 */
#include "gromacs/gpu_utils/cudautils.cuh"
#include "gromacs/gpu_utils/typecasts.cuh"

#include "nbnxm_cuda_kernel_utils.cuh"
#include "nbnxm_cuda_types.h"

/* ---- the kernels, generated the way nbnxm_cuda_kernels.cuh does it ------------------ */
#define EL_EWALD_ANA
#define NB_KERNEL_FUNC_NAME(x, ...) x##_ElecEw_VdwLJ##__VA_ARGS__
#include "nbnxm_cuda_kernel.cuh" /* F */
#define CALC_ENERGIES
#include "nbnxm_cuda_kernel.cuh" /* VF */
#undef CALC_ENERGIES
#undef NB_KERNEL_FUNC_NAME
/* ... with the Lennard-Jones force switch and potential switch (flavours _VdwLJFsw, _VdwLJPsw) */
#define LJ_FORCE_SWITCH
#define NB_KERNEL_FUNC_NAME(x, ...) x##_ElecEw_VdwLJFsw##__VA_ARGS__
#include "nbnxm_cuda_kernel.cuh"
#define CALC_ENERGIES
#include "nbnxm_cuda_kernel.cuh"
#undef CALC_ENERGIES
#undef NB_KERNEL_FUNC_NAME
#undef LJ_FORCE_SWITCH
#define LJ_POT_SWITCH
#define NB_KERNEL_FUNC_NAME(x, ...) x##_ElecEw_VdwLJPsw##__VA_ARGS__
#include "nbnxm_cuda_kernel.cuh"
#define CALC_ENERGIES
#include "nbnxm_cuda_kernel.cuh"
#undef CALC_ENERGIES
#undef NB_KERNEL_FUNC_NAME
#undef LJ_POT_SWITCH
#undef EL_EWALD_ANA

#define EL_RF
#define NB_KERNEL_FUNC_NAME(x, ...) x##_ElecRF_VdwLJ##__VA_ARGS__
#include "nbnxm_cuda_kernel.cuh"
#define CALC_ENERGIES
#include "nbnxm_cuda_kernel.cuh"
#undef CALC_ENERGIES
#undef NB_KERNEL_FUNC_NAME
#undef EL_RF

#include <cstdio>
#include <cstring>
#include <vector>

extern "C" struct nbfork_params
{
    int    eeltype;
    double epsfac, rcoulomb, rvdw, krf, crf, sh_ewald, ewaldcoeff_q, dispersion_cpot, repulsion_cpot;
    int    vdw_switch_kind; /* 0 plain, 1 force switch, 2 potential switch (Ewald flavours only) */
    double rvdw_switch;
};

#define CK(call)                                                                                   \
    do                                                                                             \
    {                                                                                              \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
        {                                                                                          \
            std::fprintf(stderr, "nbfork: %s failed: %s\n", #call, cudaGetErrorString(e_));        \
            return -2;                                                                             \
        }                                                                                          \
    } while (0)

namespace
{
template<typename T>
struct Dev
{
    T* p = nullptr;
    ~Dev() { cudaFree(p); }
    cudaError_t upload(const T* h, size_t n)
    {
        cudaError_t e = cudaMalloc(&p, sizeof(T) * (n + 1));
        if (e != cudaSuccess || n == 0)
        {
            return e;
        }
        return cudaMemcpy(p, h, sizeof(T) * n, cudaMemcpyHostToDevice);
    }
    cudaError_t zeros(size_t n)
    {
        cudaError_t e = cudaMalloc(&p, sizeof(T) * (n + 1));
        return e != cudaSuccess ? e : cudaMemset(p, 0, sizeof(T) * (n + 1));
    }
};
} // namespace

extern "C" const char* nbfork_describe()
{
    return "reference CUDA cluster-pair kernels (nbnxm_cuda_kernel.cuh) compiled in place for sm_100a";
}

/* One evaluation on zeroed outputs; ms = device time of the kernel alone (CUDA events, best of `repeats`; warm caches).
 *   xq float4[natoms] (masked charges), type int[natoms], nbfp double[2 ntype^2] = {6 C6, 12 C12}
 *   sci / cj / excl: the bytes of nbnxn_sci_t[nsci], nbnxn_cj_packed_t[ncj], nbnxn_excl_t[nexcl]
 *   f float[3 natoms], fshift float[135], e float[2] = {eLJ, eElec} */
extern "C" int nbfork_run(const nbfork_params* p, int ntype, const double* nbfp, int natoms, const float* xq, const int* type,
                          int nsci, const void* sci, int ncj, const void* cj, int nexcl, const void* excl,
                          const float* shiftvec, int energy, int repeats, float* f, float* fshift, float* e, float* ms)
{
    const bool ewald = (p->eeltype >= 3 && p->eeltype <= 6) || (p->eeltype >= 13 && p->eeltype <= 15);
    std::vector<float2> h_nbfp((size_t)ntype * ntype);
    for (size_t i = 0; i < h_nbfp.size(); i++)
    {
        h_nbfp[i] = make_float2((float)nbfp[2 * i], (float)nbfp[2 * i + 1]);
    }
    Dev<float4>            d_xq;
    Dev<int>               d_type;
    Dev<float2>            d_nbfp;
    Dev<float>             d_f, d_fshift, d_shiftvec, d_e;
    Dev<nbnxn_sci_t>       d_sci;
    Dev<nbnxn_cj_packed_t> d_cj;
    Dev<nbnxn_excl_t>      d_excl;
    CK(d_xq.upload(reinterpret_cast<const float4*>(xq), natoms));
    CK(d_type.upload(type, natoms));
    CK(d_nbfp.upload(h_nbfp.data(), h_nbfp.size()));
    CK(d_f.zeros(3 * (size_t)natoms));
    CK(d_fshift.zeros(135));
    CK(d_shiftvec.upload(shiftvec, 135));
    CK(d_e.zeros(2));
    CK(d_sci.upload(static_cast<const nbnxn_sci_t*>(sci), nsci));
    CK(d_cj.upload(static_cast<const nbnxn_cj_packed_t*>(cj), ncj));
    CK(d_excl.upload(static_cast<const nbnxn_excl_t*>(excl), nexcl));

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
    adat.xq               = d_xq.p;
    adat.f                = reinterpret_cast<Float3*>(d_f.p);
    adat.fShift           = reinterpret_cast<Float3*>(d_fshift.p);
    adat.shiftVec         = reinterpret_cast<Float3*>(d_shiftvec.p);
    adat.shiftVecUploaded = true;
    adat.eLJ              = d_e.p + 0;
    adat.eElec            = d_e.p + 1;
    adat.numTypes         = ntype;
    adat.atomTypes        = d_type.p;

    NBParamGpu nbp{};
    nbp.elecType              = ewald ? Nbnxm::ElecType::EwaldAna : Nbnxm::ElecType::RF;
    nbp.vdwType               = Nbnxm::VdwType::Cut;
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
    nbp.dispersion_shift.cpot = (float)p->dispersion_cpot;
    nbp.repulsion_shift.cpot  = (float)p->repulsion_cpot;
    if (p->vdw_switch_kind != 0)
    {
        /* force_switch_constants / potential_switch_constants, mdtypes/interaction_const.cpp:216-245, as
         * set_cutoff_parameters copies them into NBParamGpu (nbnxm_gpu_data_mgmt.cpp:201-223) */
        const double rsw = p->rvdw_switch, rc = p->rvdw, d = rc - rsw;
        auto fsw = [&](double pw, shift_consts_t* sc) {
            sc->c2 = (float)(((pw + 1) * rsw - (pw + 4) * rc) / (pow(rc, pw + 2) * d * d));
            sc->c3 = (float)(-((pw + 1) * rsw - (pw + 3) * rc) / (pow(rc, pw + 2) * d * d * d));
        };
        fsw(6.0, &nbp.dispersion_shift);
        fsw(12.0, &nbp.repulsion_shift);
        nbp.vdw_switch.c3 = (float)(-10.0 / (d * d * d));
        nbp.vdw_switch.c4 = (float)(15.0 / (d * d * d * d));
        nbp.vdw_switch.c5 = (float)(-6.0 / (d * d * d * d * d));
        nbp.rvdw_switch   = (float)rsw;
        nbp.vdwType       = p->vdw_switch_kind == 1 ? Nbnxm::VdwType::FSwitch : Nbnxm::VdwType::PSwitch;
    }
    nbp.nbfp                  = reinterpret_cast<Float2*>(d_nbfp.p);
    nbp.nbfp_texobj           = nbfp_tex;

    Nbnxm::gpu_plist pl{};
    pl.na_c      = c_clSize;
    pl.nsci      = nsci;
    pl.sci       = d_sci.p;
    pl.ncjPacked = ncj;
    pl.cjPacked  = d_cj.p;
    pl.excl      = d_excl.p;
    pl.nexcl     = nexcl;

    /* calc_shmem_required_nonbonded(), nbnxm_cuda.cu:594-623, NTHREAD_Z = 1, plain LJ */
    const int shmem = c_nbnxnGpuNumClusterPerSupercluster * c_clSize * sizeof(float4)
                      + 1 * c_nbnxnGpuClusterpairSplit * c_nbnxnGpuJgroupSize * sizeof(int)
                      + c_nbnxnGpuNumClusterPerSupercluster * c_clSize * sizeof(int);
    const dim3  block(c_clSize, c_clSize, 1);
    const dim3  grid(nsci, 1, 1);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < (repeats > 0 ? repeats : 1); r++)
    {
        CK(cudaMemset(d_f.p, 0, sizeof(float) * 3 * (size_t)natoms));
        CK(cudaMemset(d_fshift.p, 0, sizeof(float) * 135));
        CK(cudaMemset(d_e.p, 0, sizeof(float) * 2));
        CK(cudaEventRecord(e0, nullptr));
        if (nsci > 0)
        {
            if (ewald && p->vdw_switch_kind == 1)
            {
                if (energy)
                    nbnxn_kernel_ElecEw_VdwLJFsw_VF_cuda<<<grid, block, shmem>>>(adat, nbp, pl, true);
                else
                    nbnxn_kernel_ElecEw_VdwLJFsw_F_cuda<<<grid, block, shmem>>>(adat, nbp, pl, false);
            }
            else if (ewald && p->vdw_switch_kind == 2)
            {
                if (energy)
                    nbnxn_kernel_ElecEw_VdwLJPsw_VF_cuda<<<grid, block, shmem>>>(adat, nbp, pl, true);
                else
                    nbnxn_kernel_ElecEw_VdwLJPsw_F_cuda<<<grid, block, shmem>>>(adat, nbp, pl, false);
            }
            else if (ewald)
            {
                if (energy)
                    nbnxn_kernel_ElecEw_VdwLJ_VF_cuda<<<grid, block, shmem>>>(adat, nbp, pl, true);
                else
                    nbnxn_kernel_ElecEw_VdwLJ_F_cuda<<<grid, block, shmem>>>(adat, nbp, pl, false);
            }
            else
            {
                if (energy)
                    nbnxn_kernel_ElecRF_VdwLJ_VF_cuda<<<grid, block, shmem>>>(adat, nbp, pl, true);
                else
                    nbnxn_kernel_ElecRF_VdwLJ_F_cuda<<<grid, block, shmem>>>(adat, nbp, pl, false);
            }
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, nullptr));
        CK(cudaEventSynchronize(e1));
        float t = 0;
        CK(cudaEventElapsedTime(&t, e0, e1));
        best = t < best ? t : best;
    }
    CK(cudaMemcpy(f, d_f.p, sizeof(float) * 3 * (size_t)natoms, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fshift, d_fshift.p, sizeof(float) * 135, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(e, d_e.p, sizeof(float) * 2, cudaMemcpyDeviceToHost));
    *ms = best;
    cudaDestroyTextureObject(nbfp_tex);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}
