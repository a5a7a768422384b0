/*
 * fepb200_nb.h -- C-ABI of the NON-perturbed neighbour of the FEP path in libfepb200.so:
 * the cluster-pair non-bonded kernel that shares atoms, force buffer and stream with the perturbed-pair
 * kernels of fepb200.h, so that forces stay on the device, in nbat (grid) order, end to end.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   nbnxn_kernel_gpu_ref()            src/gromacs/nbnxm/kernels_reference/kernel_gpu_ref.cpp:54-354   (the arithmetic)
 *   nbnxn_atomdata_mask_fep()         src/gromacs/nbnxm/atomdata.cpp:930-964                          (masked perturbed atoms)
 *   gpu_init_atomdata()               src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:873-988
 *   gpu_init_pairlist()               src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:667-759
 *   gpu_copy_xq_to_gpu()              src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:1330-1380
 *   gpu_launch_kernel()               src/gromacs/nbnxm/cuda/nbnxm_cuda.cu:642-871  (nbnxn_kernel_*_cuda, nbnxm_cuda_kernel.cuh)
 *   gpu_launch_cpyback()              src/gromacs/nbnxm/nbnxm_gpu_data_mgmt.cpp:1117-1300
 *
 * Interactions: plain cut-off Lennard-Jones with potential shift + reaction-field or Ewald real-space
 * electrostatics (analytical), one energy group -- what the reference kernel for GPU lists computes.
 * Plain C; same error codes, flags and fepb200_params as fepb200.h.  No CPU fallback.
 */
#ifndef FEPB200_NB_H
#define FEPB200_NB_H

#include "fepb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FEPB200_NB_CLUSTER_SIZE 8          /* c_nbnxnGpuClusterSize, pairlistparams.h:65 */
#define FEPB200_NB_CLUSTERS_PER_SUPER 8    /* c_nbnxnGpuNumClusterPerSupercluster, pairlist.h:174 */
#define FEPB200_NB_JGROUP_SIZE 4           /* c_nbnxnGpuJgroupSize, pairlist.h:180 */
#define FEPB200_NB_CLUSTERPAIR_SPLIT 2     /* c_nbnxnGpuClusterpairSplit, pairlistparams.h:93 */
#define FEPB200_NB_EXCL_SIZE 32            /* c_nbnxnGpuExclSize, pairlistparams.h:97 */
#define FEPB200_NB_MIN_RSQ 3.82e-07f       /* c_nbnxnMinDistanceSquared (mixed precision), pairlist.h:167 */

/* The list structures, byte for byte the reference's (pairlist.h:195-280); a caller inside GROMACS passes
 * NbnxnPairlistGpu::sci.data(), cjPacked.list_.data(), excl.data() as they are. */
typedef struct fepb200_nb_sci
{
    int sci;           /* i-super-cluster                                  */
    int shift;         /* shift vector index                               */
    int cjPackedBegin; /* first packed j-cluster entry                     */
    int cjPackedEnd;   /* one past the last                                */
} fepb200_nb_sci;      /* nbnxn_sci_t */

typedef struct fepb200_nb_im_ei
{
    unsigned int imask;    /* bit jm*8+im: i-cluster im interacts with j-cluster jm of the entry */
    int          excl_ind; /* index into excl[] for this half of the cluster pairs; 0 = no exclusions */
} fepb200_nb_im_ei;        /* nbnxn_im_ei_t */

typedef struct fepb200_nb_cj_packed
{
    int              cj[FEPB200_NB_JGROUP_SIZE];
    fepb200_nb_im_ei imei[FEPB200_NB_CLUSTERPAIR_SPLIT];
} fepb200_nb_cj_packed; /* nbnxn_cj_packed_t */

typedef struct fepb200_nb_excl
{
    unsigned int pair[FEPB200_NB_EXCL_SIZE]; /* word (jj%4)*8+ii of half jj/4, bit jm*8+im: 1 = the atoms interact */
} fepb200_nb_excl;                           /* nbnxn_excl_t */

typedef struct fepb200_nb fepb200_nb;

int         fepb200_nb_create(fepb200_nb** h, int device_ordinal);
int         fepb200_nb_destroy(fepb200_nb* h);
const char* fepb200_nb_last_error(const fepb200_nb* h);
/* All later work of the handle goes to the caller's cudaStream_t (NULL: the handle's own), cf. fepb200_set_stream(). */
int fepb200_nb_set_stream(fepb200_nb* h, void* stream);

/* Constants: eeltype (cut-off / RF -> reaction field, PME / Ewald family -> Ewald real space, as usingFullElectrostatics),
 * epsfac, rcoulomb, rvdw, reactionFieldCoefficient, reactionFieldShift, sh_ewald, ewaldcoeff_q, dispersion/repulsion
 * shift (kernel_gpu_ref.cpp:88-101,226-287).  The soft-core fields are ignored.
 * vdw_modifier = FEPB200_MOD_FORCESWITCH or FEPB200_MOD_POTSWITCH (with rvdw_switch) selects the Lennard-Jones force / potential
 * switch the way the reference's CUDA kernels apply it (nbnxm/cuda/nbnxm_cuda_kernel_utils.cuh:104-211, constants of
 * mdtypes/interaction_const.cpp:216-245); nbnxn_kernel_gpu_ref itself knows no switch, and every other value of vdw_modifier
 * -- 0 included -- gives its arithmetic: plain LJ, the two cpot constants shifting the energy. */
int fepb200_nb_set_params(fepb200_nb* h, const fepb200_params* ic);
/* nbat->params().nbfp: {6*C6, 12*C12} for ntype x ntype types; type ntype-1 is the non-interacting type that filler atoms
 * and masked perturbed atoms carry and must have zero parameters (atomdata.cpp:930-964). */
int fepb200_nb_set_nbfp(fepb200_nb* h, int ntype, const float* nbfp);

/* ---- search-step inputs ------------------------------------------------ */
/* Atom types and charges in grid (nbat) order, UNMASKED, natoms a multiple of 8 (filler atoms: type ntype-1, charge 0);
 * replaces the type / charge part of gpu_init_atomdata(). */
int fepb200_nb_set_atoms(fepb200_nb* h, int natoms, const int* type, const float* charge);
/* nbnxn_atomdata_mask_fep() on the device copies: the listed atoms (grid indices) get type ntype-1 and charge 0 in what the
 * cluster kernel reads; their interactions belong to the perturbed-pair kernels, which keep the original parameters
 * (fepb200_set_atoms).  Undone by the next fepb200_nb_set_atoms(). */
int fepb200_nb_mask_perturbed(fepb200_nb* h, int n, const int* atoms);
/* Read the (masked) device copies back: type int[natoms], charge float[natoms] (tests). */
int fepb200_nb_get_atoms(const fepb200_nb* h, int* type, float* charge);
/* The cluster pair list; replaces gpu_init_pairlist().  Every index is range-checked (an offending list is refused). */
int fepb200_nb_set_pairlist(fepb200_nb* h, int nsci, const fepb200_nb_sci* sci, int ncj, const fepb200_nb_cj_packed* cj,
                            int nexcl, const fepb200_nb_excl* excl);

/* Read the list from the CALLER's device copies from now on (until the next fepb200_nb_set_pairlist): d_sci / d_cj / d_excl
 * must be device arrays with the layout and sizes of the host list given to fepb200_nb_set_pairlist -- the fork's
 * gpu_plist::sci / cjPacked / excl (nbnxm/gpu_types_common.h:297-340).  The work items stay those of the host list, but the
 * i-cluster masks are taken from the device copy at kernel time: what the fork's dynamic pruning kernels clear there
 * (nbnxm/cuda/nbnxm_cuda_kernel_pruneonly.cuh) is skipped here too.  NULL, NULL, NULL: back to the library's own copies. */
int fepb200_nb_use_device_list(fepb200_nb* h, const fepb200_nb_sci* d_sci, const fepb200_nb_cj_packed* d_cj,
                               const fepb200_nb_excl* d_excl);

/* ---- the hot call ------------------------------------------------------ */
#define FEPB200_NB_Q_FROM_XQ (1 << 20) /* extension bit: charges come from the .w of the caller's xyzq array */
#define FEPB200_NB_SHIFTVEC_ON_DEVICE (1 << 21) /* fepb200_nb_launch_device: shiftvec is a device pointer (NBAtomDataGpu::shiftVec) */
/* One evaluation with host buffers: x = rvec[natoms] in grid order (what nbnxn_atomdata_copy_x_to_nbat_x produces),
 * shiftvec rvec[45].  flags: FEPB200_DO_FORCE (always implied), FEPB200_DO_SHIFTFORCE, FEPB200_DO_POTENTIAL,
 * FEPB200_CLEAR_OUTPUTS.  Outputs are ACCUMULATED like the reference kernel does with clearF = enbvClearFNo:
 * f rvec[natoms], fshift rvec[45], vc[1], vvdw[1].  Host -> device copy of x and device -> host copy of f inside the call. */
int fepb200_nb_compute(fepb200_nb* h, const float* x, const float* shiftvec, int flags, float* f, float* fshift,
                       double* vc, double* vvdw);
/* The same with the coordinates as the reference holds them for its GPU-layout kernels: xq = float4[natoms], nbat->x() in
 * nbatXYZQ format (atomdata.h:77-96).  With FEPB200_NB_Q_FROM_XQ the charge is taken from .w -- the reference masks it
 * there itself -- otherwise from the masked device copy. */
int fepb200_nb_compute_xyzq(fepb200_nb* h, const float* xq, const float* shiftvec, int flags, float* f, float* fshift,
                            double* vc, double* vvdw);
/* Device-resident: d_xq = float4[natoms] (NBAtomDataGpu::xq; .w is ignored unless FEPB200_NB_Q_FROM_XQ is set -- the
 * charge normally comes from the masked device copy), d_f = float3[natoms] (NBAtomDataGpu::f) that the kernel ADDS into
 * with atomic operations -- the same buffer fepb200_add_forces_device() adds the perturbed pairs' forces into --,
 * d_fshift float[135], d_energies double[2] = {vc, vvdw}; the last three may be NULL when the flag is not set.
 * shiftvec is a HOST pointer (uploaded when it changed).  Asynchronous on the handle's stream. */
int fepb200_nb_launch_device(fepb200_nb* h, const float* d_xq, const float* shiftvec, int flags, float* d_f,
                             float* d_fshift, double* d_energies);
/* The same with the energies added (float atomics, like the fork's own kernels do) straight into the float accumulators the
 * fork's nbnxm GPU module copies back and reduces (NBAtomDataGpu::eLJ, eElec; nbnxm/gpu_types_common.h:120-122): one kernel
 * launch per step, nothing else on the stream. */
int fepb200_nb_launch_device_float_energies(fepb200_nb* h, const float* d_xq, const float* shiftvec, int flags, float* d_f,
                                            float* d_fshift, float* d_eLJ, float* d_eElec);
/* With FEPB200_DO_POTENTIAL and d_energies == NULL the launch accumulates {vc, vvdw} in a buffer of the handle's; this adds
 * them (atomically) into the float buffers the fork's nbnxm GPU module copies back and reduces (NBAtomDataGpu::eLJ, eElec;
 * nbnxm/gpu_types_common.h:120-122, gpu_common.h:139-191).  Asynchronous on the handle's stream. */
int fepb200_nb_export_energies_device(fepb200_nb* h, float* d_eLJ, float* d_eElec);
int fepb200_nb_wait(fepb200_nb* h);

/* Evidence for bench.py: kernel launches so far, device time of the last cluster kernel (CUDA events on the launching
 * stream, valid after fepb200_nb_wait), and the work it did: cluster pairs in the list (imask bits) -> 64 atom pairs each. */
long long fepb200_nb_launch_count(const fepb200_nb* h);
int       fepb200_nb_last_kernel_ms(fepb200_nb* h, float* ms);
long long fepb200_nb_cluster_pairs(const fepb200_nb* h);

#ifdef __cplusplus
}
#endif
#endif /* FEPB200_NB_H */
