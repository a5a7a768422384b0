/*
 * fepb200.h -- C-ABI of libfepb200.so, the sm_100a (B200) implementation of the
 * GROMACS free-energy-perturbation (FEP) perturbed-pair non-bonded kernel.
 *
 * This is the drop-in boundary for ONE hot path of the reference
 * (GROMACS 2023.3 + FEP-GPU fork, paths relative to the reference root):
 *
 *   gmx_nb_free_energy_kernel()            src/gromacs/gmxlib/nonbonded/nb_free_energy.h:53-72
 *   dispatchFreeEnergyKernel() + foreign-lambda loop
 *                                          src/gromacs/nbnxm/freeenergydispatch.cpp:147-308
 *   (fork GPU twins that this replaces)    src/gromacs/nbnxm/gpu_data_mgmt.h:74-108
 *                                          src/gromacs/nbnxm/cuda/nbnxm_cuda.cu:755-851
 *
 * Plain C: POD structs, raw pointers and sizes only.  No exceptions cross this
 * boundary; every entry point returns FEPB200_OK (0) or a negative error code and
 * the message is available from fepb200_last_error().  There is NO CPU fallback:
 * if no sm_100-class CUDA device is usable fepb200_create() fails.
 *
 * Threading: one host thread per context.  Each context owns one CUDA device,
 * one stream and all device memory for its shard of the pair list.
 *
 * Index space: atom indices are the caller's local topology indices (the same
 * indices t_nblist::iinr / jjnr use on the CPU path, nb_free_energy.cpp:476,546).
 */
#ifndef FEPB200_H
#define FEPB200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------- */
#define FEPB200_OK 0
#define FEPB200_ERR_INVALID_ARGUMENT (-1)
#define FEPB200_ERR_CUDA (-2)
#define FEPB200_ERR_NO_DEVICE (-3)
#define FEPB200_ERR_UNSUPPORTED (-4)
#define FEPB200_ERR_STATE (-5)

/* ---- kernel flags: identical bit values to the reference ---------------
 * src/gromacs/gmxlib/nonbonded/nonbonded.h:38-42 */
#define FEPB200_DO_FORCE (1 << 1)
#define FEPB200_DO_SHIFTFORCE (1 << 2)
#define FEPB200_DO_FOREIGNLAMBDA (1 << 3)
#define FEPB200_DO_POTENTIAL (1 << 4)
#define FEPB200_DO_SR (1 << 5)
/* extension bit (not in the reference): overwrite instead of accumulate into the
 * host output arrays of fepb200_compute() / fepb200_download().  For the force array only
 * the entries of atoms that occur in the pair list are written; all others are left alone. */
#define FEPB200_CLEAR_OUTPUTS (1 << 16)
/* extension bit for fepb200_add_forces_device() / fepb200_export_scalars_device(): add with atomic
 * operations, because kernels on ANOTHER stream may be adding into the same device buffers at the same
 * time -- the fork's local and non-local non-bonded kernels do, with atomicAdd, into one NBAtomDataGpu
 * (f, fShift, eLJ, eElec, ...; nbnxm/cuda/nbnxm_cuda_kernel_utils.cuh) when a rank has two localities. */
#define FEPB200_ATOMIC_OUTPUTS (1 << 17)

/* ---- enum values: identical integers to the reference enums ------------
 * api/legacy/include/gromacs/mdtypes/md_enums.h:238-274,324-334,640-646 */
#define FEPB200_EEL_CUT 0
#define FEPB200_EEL_RF 1
#define FEPB200_EEL_PME 3
#define FEPB200_EEL_EWALD 4
#define FEPB200_EEL_RFZERO 16
#define FEPB200_VDW_CUT 0
#define FEPB200_VDW_PME 5
#define FEPB200_MOD_POTSHIFT 1
#define FEPB200_MOD_NONE 2
#define FEPB200_MOD_POTSWITCH 3
#define FEPB200_MOD_FORCESWITCH 5
#define FEPB200_SC_BEUTLER 0
#define FEPB200_SC_GAPSYS 1

#define FEPB200_NUM_SHIFT_VECTORS 45 /* pbcutil/ishift.h:41-54 */
#define FEPB200_NUM_LAMBDA_COMPONENTS 7 /* md_enums.h:497-508 */
#define FEPB200_LAMBDA_COUL 2
#define FEPB200_LAMBDA_VDW 3

/* Mirrors, field for field, what nb_free_energy.cpp:323-396 reads from
 * interaction_const_t (mdtypes/interaction_const.h:111-184) and its
 * SoftCoreParameters (:116-138).  `real` is float (mixed-precision build). */
typedef struct fepb200_params
{
    int   eeltype;      /* CoulombInteractionType as int                       */
    int   vdwtype;      /* VanDerWaalsType as int                              */
    int   vdw_modifier; /* InteractionModifiers as int                         */
    float epsfac;
    float rcoulomb;
    float rvdw;
    float rvdw_switch;
    float reactionFieldCoefficient; /* k_rf */
    float reactionFieldShift;       /* c_rf */
    float sh_ewald;
    float sh_lj_ewald;
    float ewaldcoeff_q;
    float ewaldcoeff_lj;
    float dispersion_shift_cpot;
    float repulsion_shift_cpot;
    /* SoftCoreParameters */
    int   softcoreType; /* SoftcoreType as int */
    float alphaVdw;
    float alphaCoulomb;
    int   lambdaPower;
    float sigma6WithInvalidSigma;
    float sigma6Minimum;
    float gapsysScaleLinpointVdW;
    float gapsysScaleLinpointCoul;
    float gapsysSigma6VdW;
} fepb200_params;

typedef struct fepb200_ctx fepb200_ctx;

/* Sizes of the result block a context produces, see fepb200_result_layout(). */
typedef struct fepb200_layout
{
    int       natoms;      /* atoms in the caller's index space                 */
    int       ntouched;    /* distinct atoms that appear in the FULL list       */
    int       nri;         /* i-entries held by this context (its shard)        */
    long long nrj;         /* pairs held by this context (its shard)            */
    int       nri_total;   /* i-entries of the full list                        */
    long long nrj_total;   /* pairs of the full list                            */
    int       nenergrp;    /* energy-group pairs G                              */
    int       nforeign;    /* L (foreign lambda points, without the current one) */
    /* Device result block, identical on every rank so it can be combined across GPUs as it is.  It is ONE contiguous
     * allocation laid out [f64 part | f32 part] (fepb200_result_block_bytes() bytes; this is also the layout of a
     * published block, fepb200_publish_result() / fepb200_set_partial_result_block()); fepb200_result_device_ptrs()
     * returns the start of each part.
     *   f64 part (offsets below in units of doubles from its start):
     *     Vc[G] Vv[G] dvdl[2] foreign_E[L+1] foreign_dvdl[(L+1)][2]
     *   f32 part (fp32 words from its start):
     *     [0, 3*ntouched)          compact forces, atom k of fepb200_touched_atoms()
     *     [off_fshift, +3*45)      shift forces                                                  */
    long long f32_words;
    long long f64_words;
    long long off_fshift; /* in f32 words */
    long long off_vc, off_vv, off_dvdl, off_foreign_e, off_foreign_dvdl; /* in f64 words */
} fepb200_layout;

/* ---- lifetime ---------------------------------------------------------- */
/* Replaces Nbnxm::gpu_init(..., bFEP, n_lambda) (nbnxm_setup.cpp:463). */
int fepb200_create(fepb200_ctx** ctx, int device_ordinal);
int fepb200_destroy(fepb200_ctx* ctx);
/* Message of the last failing call on this context ("" if none).  ctx may be
 * NULL to read the message of a failed fepb200_create(). */
const char* fepb200_last_error(const fepb200_ctx* ctx);
/* Library/device description, e.g. "fepb200 0.1 sm_100a NVIDIA B200 148 SMs". */
const char* fepb200_describe(const fepb200_ctx* ctx);

/* Makes every later operation of the context (copies, kernels, timing events) use the
 * caller's cudaStream_t instead of the context's own stream, the way the reference hands its
 * per-locality DeviceStream to the nbnxm GPU module (nbnxm_gpu_data_mgmt.cpp:404-423).
 * NULL restores the context's own stream.  The context's previous stream is drained first. */
int fepb200_set_stream(fepb200_ctx* ctx, void* stream);

/* ---- constants (init time) --------------------------------------------- */
/* Replaces cuda_copy_fepparams() (nbnxm/gpu_data_mgmt.h:74-85) and the reads at
 * nb_free_energy.cpp:323-396. */
int fepb200_set_params(fepb200_ctx* ctx, const fepb200_params* params);
/* nbfp = {6*C6, 12*C12} interleaved, real[2*ntype*ntype] (mdlib/forcerec.cpp:115-152);
 * nbfp_grid likewise (only [2k] used; forcerec.cpp:154-190), may be NULL unless LJ-PME. */
int fepb200_set_nbfp(fepb200_ctx* ctx, int ntype, const float* nbfp, const float* nbfp_grid);

/* ---- search-step inputs ------------------------------------------------ */
/* chargeA/B, typeA/B of nb_free_energy.h:61-64 (replaces setAtomPropertiesAB,
 * atomdata.cpp:1055-1072).  Must precede fepb200_set_list(). */
int fepb200_set_atoms(fepb200_ctx* ctx, int natoms, const float* chargeA, const float* chargeB,
                      const int* typeA, const int* typeB);
/* The FEP t_nblist (mdtypes/nblist.h:41-55), replaces gpu_init_feppairlist()
 * (nbnxm_gpu_data_mgmt.cpp:761-871).  excl_fep may be NULL (= all included,
 * nb_free_energy.cpp:545).  nenergrp_pairs = G, every gid must be < G.
 * rank/nranks: the context keeps the rank-th of nranks contiguous i-entry ranges,
 * balanced by pair count with the rule of balance_fep_lists() (pairlist.cpp:2786-2838);
 * use rank=0,nranks=1 for a single GPU. */
int fepb200_set_list(fepb200_ctx* ctx, int nri, const int* iinr, const int* gid, const int* shift,
                     const int* jindex, const int* jjnr, const int* excl_fep, int nenergrp_pairs,
                     int rank, int nranks);
/* The same from the lists as the reference holds them: one t_nblist per OpenMP thread and locality
 * (nbnxm/pairlistset.h fepLists()), optionally in another index space than fepb200_set_atoms().  Replaces what the
 * fork does on the HOST before its GPU FEP kernels can run -- combine_fep_lists() (nbnxm/pairlist.cpp:2867-2961:
 * element-wise concatenation of the per-thread lists) and the index remap loops of gpu_init_feppairlist()
 * (nbnxm_gpu_data_mgmt.cpp:763-787: local atom index -> nbat index through the inverse of nbat->cell) -- by
 * n_lists bulk copies straight into the concatenated device arrays and three small kernels (jindex offsets,
 * remap, range check).  The list the context holds afterwards (fepb200_get_list) is the concatenation in list
 * order, in the index space of fepb200_set_atoms().
 *   lists[l]   as in fepb200_set_list(); excl_fep must be given for all lists or for none
 *   atom_map   NULL, or int[n_map]: atom index used by the lists -> atom index of fepb200_set_atoms()
 *              (the GPU route passes the fork's atomIndicesInv here) */
typedef struct fepb200_list_view
{
    int        nri;
    const int* iinr;
    const int* gid;
    const int* shift;
    const int* jindex; /* [nri + 1], jindex[0] == 0 */
    const int* jjnr;
    const int* excl_fep;
} fepb200_list_view;
int fepb200_set_lists(fepb200_ctx* ctx, int n_lists, const fepb200_list_view* lists, const int* atom_map, int n_map,
                      int nenergrp_pairs, int rank, int nranks);
/* Round trip of the shard this context holds, bit-exact (tests).
 * Call with NULL arrays to query sizes through fepb200_result_layout(). */
int fepb200_get_list(const fepb200_ctx* ctx, int* first_entry, int* iinr, int* gid, int* shift,
                     int* jindex, int* jjnr, int* excl_fep);
/* Global indices of the touched atoms, int[ntouched], ascending. */
int fepb200_touched_atoms(const fepb200_ctx* ctx, int* atoms);
int fepb200_result_layout(const fepb200_ctx* ctx, fepb200_layout* layout);

/* ---- per-step inputs --------------------------------------------------- */
/* lambda[7] as passed to gmx_nb_free_energy_kernel (only [COUL],[VDW] are read,
 * nb_free_energy.cpp:319-320).  n_foreign = L with all_lambda_coul/vdw[L] =
 * fepvals->all_lambda[Coul|Vdw][0..L) (freeenergydispatch.cpp:247-253); L may be 0. */
int fepb200_set_lambdas(fepb200_ctx* ctx, const float* lambda, int n_foreign,
                        const float* all_lambda_coul, const float* all_lambda_vdw);

/* ---- the hot call ------------------------------------------------------ */
/* One force/energy evaluation of the perturbed pairs this context holds, i.e. what
 * dispatchFreeEnergyKernel() does per step (freeenergydispatch.cpp:147-308):
 *   - the pass at the current lambda honouring DO_FORCE / DO_SHIFTFORCE / DO_POTENTIAL;
 *   - if DO_FOREIGNLAMBDA is set and L+1 > 0: the energy-only passes i = 0..L
 *     (i = 0 repeats the current lambda) returning
 *     foreign_energy[i] = sum_g Vc[g]+Vv[g] and foreign_dvdl[i][{coul,vdw}].
 * Host buffers; x is rvec[natoms] (AoS xyz), shiftvec rvec[45].  Outputs are
 * ACCUMULATED (+=) exactly like the reference kernel does into its thread buffers
 * (nb_free_energy.cpp:1155-1178) unless FEPB200_CLEAR_OUTPUTS is set.  Any output
 * pointer may be NULL when the corresponding flag is not set.
 * Copies host->device (touched coordinates only) inside the call; the results travel back as
 * they are produced: the last kernel writes them into the library's pinned host buffer over PCIe
 * (no separate device->host copy), from where they are added into the caller's arrays.  The device
 * result block (fepb200_result_device_ptrs) is only filled by fepb200_launch(). */
int fepb200_compute(fepb200_ctx* ctx, const float* x, const float* shiftvec, int flags, float* f,
                    float* fshift, double* Vc, double* Vv, double* dvdl /*[2]: coul, vdw*/,
                    double* foreign_energy /*[L+1]*/, double* foreign_dvdl /*[L+1][2]*/);

/* ---- device-resident variants (replace gpu_launch_kernel / gpu_launch_cpyback /
 * gpu_wait_finish_task, nbnxm_cuda.cu:642, nbnxm_gpu_data_mgmt.cpp:1117, gpu_common.h:405) */
/* Stage coordinates: host rvec[natoms] -> device (touched atoms only). */
int fepb200_upload_x(fepb200_ctx* ctx, const float* x, const float* shiftvec);
/* Or gather them on the device from a device-resident rvec[natoms] array. */
int fepb200_gather_x_device(fepb200_ctx* ctx, const float* d_x, const float* shiftvec);
/* The same from the xyzq array of the nbnxm GPU atom data, float4[natoms] with the charge in .w
 * (NBAtomDataGpu::xq, nbnxm/gpu_types_common.h:103-157, which the fork's FEP kernels read): with a
 * pair list in nbat (grid) indices -- the index space the fork uses after its inverse map,
 * nbnxm_gpu_data_mgmt.cpp:763-787 -- coordinates never leave the device.  The .w component is
 * ignored: charges of both states come from fepb200_set_atoms(). */
int fepb200_gather_xq_device(fepb200_ctx* ctx, const float* d_xq, const float* shiftvec);
/* Launch all kernels of one step on `stream` (a cudaStream_t; NULL = the context's
 * stream).  Results stay in the device result block.  Asynchronous. */
int fepb200_launch(fepb200_ctx* ctx, int flags, void* stream);
/* Device-resident hand-off of the forces (replaces what the fork does with atomic adds
 * into its nbat force buffer followed by nbnxn_gpu_add_nbat_f_to_f, nbnxm/atomdata.cpp:930-964, and
 * the host scatter at the end of fepb200_download()): adds the forces of the last fepb200_launch()
 * into d_f, a device-resident rvec[natoms] array in the caller's index space (the space of d_x of
 * fepb200_gather_x_device), on the context's stream; with FEPB200_CLEAR_OUTPUTS in flags the
 * entries of the atoms that occur in the pair list are overwritten instead.  Asynchronous.  After
 * fepb200_reduce_peers() the sum over ranks is added; with the fused exchange the forces of the
 * atoms this rank owns.  Energies, dV/dlambda and shift forces still come from fepb200_download()
 * (call it with FEPB200_DO_FORCE cleared to skip the force copy). */
int fepb200_add_forces_device(fepb200_ctx* ctx, float* d_f, int flags);
/* The scalars of the last fepb200_launch() added into the float device buffers the fork's nbnxm GPU module
 * copies back and reduces (NBAtomDataGpu::eLJ, eElec, dvdlLJ, dvdlElec, e{LJ,Elec}Foreign[L+1],
 * dvdl{LJ,Elec}Foreign[L+1], fShift[45]; nbnxm/gpu_types_common.h:103-157, reduced at gpu_common.h:139-191), on
 * the context's stream -- what a hook in the fork's GPU route (nbnxm/cuda/nbnxm_cuda.cu:755-851) needs so that
 * the fork's copy-back and reduction stay as they are.  flags select the groups as in fepb200_download()
 * (DO_POTENTIAL: eLJ / eElec, summed over energy-group pairs, which the fork does not have; DO_FOREIGNLAMBDA:
 * the foreign arrays, the energy of a point whole in eLJForeign; DO_FORCE|DO_SHIFTFORCE: fShift; dV/dlambda
 * always).  Any pointer may be NULL.  Asynchronous. */
int fepb200_export_scalars_device(fepb200_ctx* ctx, int flags, float* eLJ, float* eElec, float* dvdlLJ, float* dvdlElec,
                                  float* eLJForeign, float* eElecForeign, float* dvdlLJForeign, float* dvdlElecForeign,
                                  float* fShift);
/* Block until the context's stream is idle. */
int fepb200_wait(fepb200_ctx* ctx);
/* Device pointers of the result block: f32 part and f64 part (see fepb200_layout). */
int fepb200_result_device_ptrs(const fepb200_ctx* ctx, void** d_f32, void** d_f64);
/* ---- multi-GPU reduction over peer memory (NVLink), the alternative to handing the device
 * pointers to ncclAllReduce.  Every rank (1) publishes its result block into a buffer that all
 * ranks of the node have mapped (e.g. CUDA VMM / symmetric memory; fepb200_result_block_bytes()
 * bytes, 16-byte aligned), then (2) calls fepb200_reduce_peers() with the device pointers of ALL
 * ranks' buffers in rank order: one kernel reads every block over NVLink and leaves the full sum
 * in this context's own result block (ready for fepb200_download()).  Sums are taken in rank
 * order: every rank gets bit-identical results.
 * The cross-GPU barrier between (1) and (2) is part of the same kernel when d_peer_flags is given:
 * per rank a peer-mapped, zero-initialised array of 16 uint32; `seq` must be the same on all
 * ranks and increase by one per step.  With d_peer_flags == NULL the caller provides the barrier
 * on the context's stream.  Use two alternating buffers so that a fast rank cannot overwrite a
 * block a slow rank is still reading. */
size_t fepb200_result_block_bytes(const fepb200_ctx* ctx);
int    fepb200_publish_result(fepb200_ctx* ctx, void* d_block);
/* Instead of publishing with a copy: make the following fepb200_launch() calls write this rank's
 * (partial) result block directly at d_block (NULL = back to the context's own block).
 * fepb200_reduce_peers() always writes the sum into the context's own block. */
int    fepb200_set_partial_result_block(fepb200_ctx* ctx, void* d_block);
int    fepb200_reduce_peers(fepb200_ctx* ctx, int nranks, void* const* d_peer_blocks, void* const* d_peer_flags,
                            int rank, unsigned int seq);
/* The same with the split BASELINE.json's north_star names: a force REDUCE-SCATTER plus an all-reduce of what is small.
 * Afterwards this context's result block holds the forces of the atoms it OWNS -- the compact atoms
 * [atom_begin, atom_end) of fepb200_peer_ranges(): equal ranges of the full list's touched atoms, starting on
 * multiples of four -- summed over all ranks (zeros elsewhere: the sum over ranks of what fepb200_download() /
 * fepb200_add_forces_device() deliver is the full force array), and the shift forces, Vc/Vv, dV/dlambda and foreign
 * terms summed over all ranks on every rank.  (N-1)/N of ONE block crosses NVLink per rank instead of N-1 blocks, and
 * only a rank's own atoms cross PCIe afterwards.  Same arguments, barrier and slot discipline as fepb200_reduce_peers();
 * replaces ThreadedForceBuffer::reduce (mdtypes/threaded_force_buffer.cpp:320-402) across GPUs. */
int    fepb200_reduce_scatter_peers(fepb200_ctx* ctx, int nranks, void* const* d_peer_blocks, void* const* d_peer_flags,
                                    int rank, unsigned int seq);
/* The same reduction with ONE one-way NVLink trip per step instead of an announcement plus a pull: the kernel that forms
 * this rank's sums stores them straight into the ranks that need them.  Every rank r owns, per step slot, `nranks`
 * zero-initialised receive blocks of fepb200_result_block_bytes() bytes, block s written by rank s only.
 * d_peer_blocks[r] = THIS rank's receive block on rank r (peer-mapped address).  The following fepb200_launch() calls
 * store the force of a compact atom into the block on the rank that owns the atom (the ranges of
 * fepb200_reduce_scatter_peers()) and the shift forces and scalars into the blocks on all ranks; nothing is written to a
 * block of this context's own.  Follow each launch with fepb200_reduce_scatter_peers(ctx, nranks, LOCAL receive blocks
 * in rank order, flags, rank, seq): its barrier orders the pushes before the sums, and the sums read local memory only.
 * The set of words a rank writes is fixed by its list: the blocks must be zeroed again (on all ranks, behind a barrier)
 * after a new fepb200_set_list(), which also switches the push off.  nranks <= 1 or d_peer_blocks == NULL: off.  At most 8
 * ranks (the targets travel as kernel arguments).  Measured on C5: 46.9 against 48.5 us per step on 2 B200, 39.3 against
 * 44.0 us on 4. */
int    fepb200_set_push_targets(fepb200_ctx* ctx, int nranks, void* const* d_peer_blocks);

/* ---- multi-GPU, fused: no separate collective, every rank sums its own atoms ------------------
 * The force reduce-scatter and the scalar all-reduce of SURVEY 8e inside the epilogue kernel
 * (replaces what the reference does with per-thread buffers + ThreadedForceBuffer::reduce,
 * threaded_force_buffer.cpp:320-402, across GPUs instead of across OpenMP threads):
 *   1. every rank calls fepb200_set_list() with the FULL list (rank 0 of 1), so that all ranks
 *      hold the same atom-sorted slot layout;
 *   2. every rank allocates fepb200_exchange_bytes() bytes of zero-initialised memory that all
 *      ranks of the node have mapped (CUDA VMM / symmetric memory), and calls
 *      fepb200_set_peer_exchange() with the device pointers of ALL ranks' buffers in rank order;
 *   3. from then on fepb200_launch() / fepb200_compute() of rank r evaluates the r-th share of the
 *      32-pair warps of the flat pair space with the unchanged pair kernels, which leave their
 *      results (force contributions at their atom-sorted slots, segment shift forces / energies,
 *      per-CTA dV/dlambda and foreign-energy partials) in rank r's own exchange buffer.  The
 *      epilogue first passes a cross-GPU barrier (sequence flags in the exchange buffers), then
 *      every rank reads over NVLink, from whichever rank produced each element, the contributions
 *      of the atoms it owns (contiguous atom ranges with equal numbers of contributions; entries
 *      of all other atoms in the result block stay zero) and ALL scalar inputs (shift forces,
 *      Vc/Vv, dV/dlambda, foreign energies: summed identically on every rank).  Forces, shift
 *      forces and Vc/Vv are bit-identical to the single-GPU result (same slots, same summation
 *      order).
 * All ranks must call fepb200_launch()/fepb200_compute() the same number of times (lockstep); a
 * rank that waits more than 4 s for a peer traps.  fepb200_download()/fepb200_compute() add the
 * owned atoms' forces and the full scalars into the caller's arrays: sum the force arrays over
 * ranks (or keep them distributed), take the scalars from one rank.
 * nranks == 1 (d_peer_bufs may be NULL) switches the exchange off.  fepb200_set_list() switches it
 * off as well: call fepb200_set_peer_exchange() again after every search step. */
size_t fepb200_exchange_bytes(const fepb200_ctx* ctx, int nranks);
int    fepb200_set_peer_exchange(fepb200_ctx* ctx, int nranks, int rank, void* const* d_peer_bufs, size_t bytes);
/* The trips [pair_begin, pair_end) (groups of <= 32 pairs that share an owner atom, the unit of the device
 * layout) this context evaluates and the compact
 * atoms [atom_begin, atom_end) (indices into fepb200_touched_atoms()) it owns; any pointer may be NULL. */
int    fepb200_peer_ranges(const fepb200_ctx* ctx, int* pair_begin, int* pair_end, int* atom_begin, int* atom_end);

/* Diagnosis of the exchange: with enable != 0 the epilogue of every following launch records, per
 * block, the GPU's global timer (ns) at block entry, when this rank's pair kernels had completed,
 * when the cross-GPU barrier had been passed, and when the block's sums were done.  stamps (may be
 * NULL) receives 4 values per block of the last launch for min(max_blocks, 4096) blocks; returns
 * that number of blocks (>= 0) or a negative error code.  Blocks are ordered reduction jobs,
 * scalar sums, per-atom sums. */
int fepb200_epilogue_trace(fepb200_ctx* ctx, int enable, unsigned long long* stamps, int max_blocks);

/* Copy the result block to the host and add it into the caller's arrays (same
 * semantics as the tail of fepb200_compute).  Synchronous. */
int fepb200_download(fepb200_ctx* ctx, int flags, float* f, float* fshift, double* Vc, double* Vv,
                     double* dvdl, double* foreign_energy, double* foreign_dvdl);

/* Number of kernel launches issued by this context since creation (bench evidence),
 * and device time in ms of the last fepb200_launch() measured with CUDA events on
 * the launching stream (valid after fepb200_wait()). */
long long fepb200_launch_count(const fepb200_ctx* ctx);
int       fepb200_last_launch_ms(fepb200_ctx* ctx, float* ms);
/* Profiling mode (replaces the reference's GpuRegionTimer fep_k, gpu_types_common.h:275-276):
 * when on, fepb200_launch() also records CUDA events between its kernels and
 * fepb200_kernel_ms() returns ms[3] = device time of {current-lambda pass kernel,
 * foreign-lambda kernel, epilogue kernel} of the last launch. */
int fepb200_set_profiling(fepb200_ctx* ctx, int on);
int fepb200_kernel_ms(fepb200_ctx* ctx, float* ms);

/* ---- perturbed 1-4 pair interactions (the step next to the path) -----------
 * Replaces the perturbed (bFreeEnergy) branch of do_pairs(F_LJ14, ...)
 * (src/gromacs/listed_forces/pairs.cpp:516-835, free_energy_evaluate_single :170-515) and the
 * fork's pairs_fep_gpu (listed_forces/listed_forces_gpu_internal.cu:1365-1500): plain Coulomb
 * (scaled by fudgeQQ) and plain LJ with the Beutler or Gapsys soft-core, no cut-off.
 * Only PERTURBED pairs belong here; the others stay with the reference's tabulated evaluation.
 * The pairs run through the same sm_100a kernels as the non-bonded perturbed pairs. */
typedef struct fepb200_pairs14 fepb200_pairs14;
int         fepb200_pairs14_create(fepb200_pairs14** h, int device_ordinal);
int         fepb200_pairs14_destroy(fepb200_pairs14* h);
const char* fepb200_pairs14_last_error(const fepb200_pairs14* h);
/* ic: epsfac, rcoulomb (only the Gapsys linearisation point looks at it) and the
 * SoftCoreParameters fields are used; fudgeQQ = fr->fudgeQQ (pairs.cpp:624). */
int fepb200_pairs14_set_params(fepb200_pairs14* h, const fepb200_params* ic, float fudgeQQ);
/* iatoms: int[3*npairs] = {1-4 type, ai, aj} (the t_iatom layout of the reference);
 * c6A/c12A/c6B/c12B: t_iparams::lj14 per 1-4 type; gid: energy-group pair per pair (may be NULL). */
int fepb200_pairs14_set_pairs(fepb200_pairs14* h, int natoms, const float* chargeA, const float* chargeB, int npairs,
                              const int* iatoms, int ntypes, const float* c6A, const float* c12A, const float* c6B,
                              const float* c12B, const int* gid, int nenergrp_pairs);
/* x rvec[natoms]; box_diag float[3] (rectangular box); pbc_type 0 = none (molecules whole),
 * 1 = xyz, 2 = xy (pbc_dx_aiuc, pbcutil/pbc.cpp:825-851); flags: FEPB200_DO_FORCE,
 * _SHIFTFORCE, _POTENTIAL.  Outputs are accumulated: f rvec[natoms], fshift rvec[45],
 * Vc14/Vv14[G] (Coulomb-14 / LJ-14 energy terms), dvdl[2] (coul, vdw). */
int fepb200_pairs14_compute(fepb200_pairs14* h, const float* x, const float* box_diag, int pbc_type, const float* lambda,
                            int flags, float* f, float* fshift, double* Vc14, double* Vv14, double* dvdl);
/* The energy-only evaluations at ALL foreign lambda points in one call: replaces the n_points calls of the pair code that
 * the reference makes from calc_listed_lambda (listed_forces/listed_forces.cpp:554-640, called per point at :760-800).
 * lambda_coul / lambda_vdw: float[n_points] = fepvals->all_lambda[Coul | Vdw]; energy: double[n_points], Coulomb-14 +
 * LJ-14 at each point summed over the energy-group pairs; dvdl: double[2 n_points] = {coul, vdw} per point.  Outputs are
 * STORED.  The pairs go through the foreign-lambda machinery of the non-bonded kernels (one load of every pair). */
int fepb200_pairs14_compute_foreign(fepb200_pairs14* h, const float* x, const float* box_diag, int pbc_type, int n_points,
                                    const float* lambda_coul, const float* lambda_vdw, double* energy, double* dvdl);

#ifdef __cplusplus
}
#endif
#endif /* FEPB200_H */
