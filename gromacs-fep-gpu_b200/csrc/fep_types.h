/*
 * fep_types.h -- data layout shared by the host side (fepb200_api.cu) and the sm_100a kernels.
 * Vocabulary follows the reference: i-entries, j atoms, pairs, energy-group pairs (gid), shift
 * vectors, lambda states A/B.
 *
 * The t_nblist handed over at a search step is regrouped ON THE DEVICE (fep_list_build.cu) into
 * TRIPS: a trip is the work of one warp at one time, 32 pair slots that share
 *   - an OWNER atom: the end of the pair that occurs in more pairs of the list (for an FEP list:
 *     the perturbed atom, whichever side of the reference's half list it sits on),
 *   - the energy-group pair, the shift vector, and the orientation (owner was i / owner was j).
 * Everything that the reference looks up per i-entry (nb_free_energy.cpp:466-503) is therefore
 * warp-uniform.  A pair whose owner was the reference's j atom is evaluated with the negated shift
 * vector (index 44 - s), which gives the negated distance vector: same energies, same forces on both
 * atoms; its contribution to the shift force is booked with the opposite sign under the original
 * shift index.
 *
 * RUNS and SEGMENTS.  The trips are cut into runs of `run_trips` consecutive trips (a power of two,
 * chosen per list so that one wave of resident warps covers it); a warp evaluates whole runs, one
 * trip after the other.  A SEGMENT is a maximal sequence of trips of one group (owner, energy-group
 * pair, shift, orientation) inside a run.  The warp keeps the owner's force (and the segment's
 * Vc/Vv) in per-lane accumulators over the segment and reduces them ONCE, at the segment's last
 * trip -- C5: 30 944 trips, 4 712 groups, 8 005 segments at run_trips = 8 -- so the owner's sum
 * costs a quarter of a warp reduction per trip, and the atom-sorted buffer holds one contribution
 * per segment instead of one per trip.
 *
 * Device layout of one context ("compact" = index into the ascending list of atoms that occur
 * anywhere in the FULL list, so that every rank uses the same numbering; "slot" = 32 * trip + lane):
 *
 *   dyn          DynHead           per step: shift vectors + constants of the current lambda
 *   pts[L+1]     LambdaPoint       per set_lambdas: point 0 = current lambda, 1.. = foreign
 *   pos3[nT]     float[3] {x,y,z}  per step: coordinates of the touched atoms (compact order),
 *                                  packed: 12 bytes per atom cross PCIe, not 16
 *   par4[nT]     float4 {qA,qB,bits(typeA),bits(typeB)}      per search step (list build only)
 *   typetab[T*T] float4 {c6,c12,sigma6,c6grid}               per nbfp upload
 *   trips[NT]    one contiguous block of FEP_TRIP_WORDS 32-bit words per trip, so that a trip is ONE
 *                bulk copy into shared memory and every field is base + constant offset:
 *                  header (16 words): owner | shift_eff << 24 | flipped << 30; flags (last trip of
 *                    its segment); the segment's slots in fsorted / fshift_sorted / ev2 (valid in
 *                    the segment's last trip); owner charges qA, qB; padding
 *                  cjx[32]  compact partner | excluded << 31 | padding << 30
 *                  dst[32]  where the force on the partner goes in fsorted
 *                  qA[32], qB[32]  partner charges       (pre-gathered: the only dependent per-pair
 *                  tj[32]   type-table index of the pair  loads left are the partner's coordinates
 *                           in state A | state B << 16     and the two type-table rows)
 *   orig[32 NT]  index of the pair in the shard's t_nblist; tgid[NT] (list read-back only)
 *   ent4[E]      int4   {compact i, shift index, gid, 0}  (list read-back only) per search step
 *   fsorted[P+NS] float4 force contributions SORTED BY RECEIVING ATOM: atom k owns the contiguous
 *                range [atom_ptr[k], atom_ptr[k+1]); the pass kernel scatters -f (pairs, to the
 *                partner) and the segment's sum (to the owner) to precomputed unique slots, the
 *                epilogue streams ranges: no atomics, bit-reproducible
 *   fshift_sorted[NS] float4 segment sums sorted by shift index;  ev2[NS] float2 segment {Vc,Vv}
 *                sorted by energy-group pair; red_jobs = chunks of those ranges
 *   cta_part[4][nPart]       fp64 partial dV/dlambda (and Vc, Vv) of the current-lambda pass
 *   for_part[3*(L+1)][nTile] fp64 per-CTA partial foreign energies / dV/dlambda
 *   result block: res_f32[3*nT + 3*45], res_f64[2G + 2 + 3(L+1)]   (include/fepb200.h)
 */
#ifndef FEPB200_FEP_TYPES_H
#define FEPB200_FEP_TYPES_H

#include <cuda_runtime.h>
#ifdef __cplusplus
#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>
#endif

#define FEP_NUM_SHIFT 45
#define FEP_MAX_TOUCHED (1 << 24) /* compact atom index, shift index and flags share one word */
#define FEP_SLOT_PADDING 0x40000000 /* cjx: the slot holds no pair (tail of a trip) */
/* layout of a trip block, in 32-bit words */
#define FEP_TRIP_WORDS 176
#define FEP_TH_OWNER 0
#define FEP_TH_SLOT_F 1
#define FEP_TH_SLOT_SHIFT 2
#define FEP_TH_SLOT_EV 3
#define FEP_TH_FLAGS 4 /* FEP_TRIP_LAST: last trip of its segment (reduce and store the owner's sums) */
#define FEP_TH_QA 6
#define FEP_TH_QB 7
#define FEP_TW_CJX 16
#define FEP_TW_DST 48
#define FEP_TW_QA 80
#define FEP_TW_QB 112
#define FEP_TW_TJ 144
#define FEP_TRIP_FLIPPED 0x40000000 /* header word 0: the owner was the j atom of the reference's pairs */
#define FEP_TRIP_LAST 1u
#define FEP_MAX_RUN_TRIPS 8 /* run_trips is a power of two <= this */
#define FEP_MAX_NTYPE 256   /* both type-table indices of a pair share one word */
#define FEP_CENTRAL_SHIFT 22
#define FEP_MAX_POINTS 256 /* L+1 <= 256 lambda points per step */
#define FEP_CTA 256        /* threads per CTA of the pair kernels */
#define FEP_LCHUNK 8       /* lambda points evaluated per thread per pair in the foreign kernel */
#define FEP_RED_CHUNK 2048 /* segments per reduction job of the epilogue */
#define FEP_EPI_CTA 256
#define FEP_EPI_LANES 8 /* lanes per touched atom in the epilogue (32 contributions per trip; 4 lanes measured slower) */
#define FEP_HEAVY_MIN 32 /* atoms with more force contributions than this are summed by a whole warp */
#define FEP_HEAVY_LOADS 8 /* independent 16-byte loads per lane and trip of that warp (256 contributions per trip) */

/* soft-core flavour actually evaluated (nb_free_energy.cpp:1324-1363) */
enum { FEP_SC_NONE = 0, FEP_SC_BEUTLER = 1, FEP_SC_GAPSYS = 2 };

/* Everything about one lambda point the kernels need (nb_free_energy.cpp:420-449). 96 bytes. */
struct LambdaPoint
{
    float lfac_c[2], lfac_v[2];     /* {1-lambda, lambda} */
    float sclfac_c[2], sclfac_v[2]; /* soft-core lambda factors */
    float scdl_c[2], scdl_v[2];     /* dlfac * p/6 * (p==2 ? 1-lfac : 1) */
    float g6_c[2], g6_v[2];         /* Gapsys: (1-lfac)^(1/6) */
    float gdl_c[2], gdl_v[2];       /* Gapsys: lfac/(1-lfac) (0 when lfac == 1) */
    int   differ;                   /* scLambdasOrAlphasDiffer for this point (:1405-1419) */
    int   pad[3];
};

/* Per-step head block, uploaded in one copy together with the coordinates that follow it
 * in device memory (step_in = [DynHead | pos3[nT]]). */
struct DynHead
{
    float4      shiftvec[FEP_NUM_SHIFT];
    LambdaPoint cur; /* the current lambda (pass at the current lambda) */
};

/* One reduction job of the epilogue: elements [begin,end) of fshift_sorted (kind 0) or ev2
 * (kind 1) all belong to output `key`. */
struct RedJob
{
    int begin, end;
    int key;  /* shift index (kind 0) or gid (kind 1) */
    int kind; /* 0: fshift_sorted -> shift force, 1: ev2 -> Vc/Vv */
};

/* Multi-GPU "owner computes" exchange (fepb200_set_peer_exchange): every rank holds the layout of
 * the FULL list and evaluates a contiguous range of its trips with the
 * unchanged single-GPU pair kernels; they store force contributions, trip sums and per-CTA
 * partials at their usual places in the rank's OWN exchange slot, which every rank of the node has
 * mapped (NVLink peer memory).  Nothing is sent: after a cross-GPU barrier at the top of the
 * epilogue each rank PULLS what it needs -- the contributions of the atoms it owns (a contiguous
 * atom range per rank: the forces are reduce-scattered) and all scalar inputs (all-reduced: every
 * rank sums them in the same order) -- from whichever rank produced each element.  The producer of
 * every element is static (it follows from the pair ranges) and kept in byte tables, and because
 * the sorted orders are stable in the pair index, consecutive elements mostly share a producer,
 * so the remote reads are coalesced.  nranks <= 1: off. */
#define FEP_XMAX 8
#define FEP_TRACE_BLOCKS 4096
struct PeerExchange
{
    int          nranks, rank;
    int          atom_begin, atom_end;   /* compact atoms whose forces this rank sums */
    int          heavy_begin, heavy_end; /* the part of heavy_atoms[] inside that range */
    int          light_begin, light_end; /* the part of light_atoms[] inside that range */
    unsigned int seq;                    /* step number announced in the barrier (set per launch) */
    int          pad;
    /* producer rank of every element of fsorted / fshift_sorted / ev2 (local, static per list) */
    const unsigned char* slot_src;
    const unsigned char* fshift_src;
    const unsigned char* ev2_src;
    /* the exchange slot of this step on every rank (index = rank) */
    const float4* fsorted[FEP_XMAX];
    const float4* fshift_sorted[FEP_XMAX];
    const float2* ev2[FEP_XMAX];
    const double* cta_part[FEP_XMAX];
    const double* for_part[FEP_XMAX];
    unsigned int* flags[FEP_XMAX]; /* per rank FEP_XMAX sequence numbers, slot r written by rank r */
};

/* Multi-GPU "push" reduction (fepb200_set_push_targets): the epilogue stores what it has summed straight into the memory
 * of the ranks that need it, over NVLink -- the force of compact atom a into the receive block of the rank that owns a
 * (equal ranges of `per_rank` atoms), the shift forces and scalars into the receive blocks of all ranks -- so that the
 * reduction kernel behind the cross-GPU barrier (fep_peer_reduce_scatter_kernel on the rank's own receive blocks) reads
 * local memory only: one one-way NVLink trip per step instead of an announcement plus a pull (a round trip and a half).
 * nranks <= 1: the epilogue writes res_f32 / res_f64 as usual. */
struct PushTargets
{
    int     nranks, per_rank;
    float*  f32[FEP_XMAX]; /* this rank's receive block on rank r: f32 part ... */
    double* f64[FEP_XMAX]; /* ... and f64 part */
};

/* Static constants + device pointers, passed by value as the kernel parameter. */
struct KernelArgs
{
    /* interaction constants (nb_free_energy.cpp:323-396) */
    float epsfac, rcoulomb, rvdw, rvdw_switch, krf, crf, sh_ewald, sh_lj_ewald;
    float beta, beta2, beta3, lj_coeff_sq, lj_coeff6_div6, disp_cpot, rep_cpot;
    float rcut_max2, rcoulomb6, rvdw6;
    float alpha_c, alpha_v, gscale_c, gscale_v;
    float gapsys_facel, gapsys_rcoul; /* epsfac and r_coulomb as seen by the Gapsys Coulomb linearisation */
    float sw_v3, sw_v4, sw_v5, sw_f2, sw_f3, sw_f4;
    int   vdw_ewald, pot_switch, rf_type, ntype;
    /* sizes */
    int n_pairs, n_entries, n_trips, n_segs, n_touched, n_gid, n_cta, n_tiles;
    int tile_trips;           /* generic foreign kernel: trips per CTA */
    int run_trips;            /* trips per run (power of two) */
    int trip_begin, trip_end; /* the trips this context evaluates ([0, n_trips) unless the list is split over peers);
                                 trip_begin is a multiple of run_trips */
    int n_points, n_chunks, chunk_points;
    int pass_n_tiles; /* CTAs of the force-only Beutler kernel */
    int n_parts;      /* partial dV/dlambda sums written by the pass of this step */
    int fuse_pass_and_foreign;         /* Beutler path: pass + first foreign chunk in one launch */
    int gapsys_hoisted;                /* Gapsys without potential switch: foreign passes by fep_gapsys.cu */
    int n_red_jobs, n_shift_jobs;
    int pdl_chain;                     /* set per step by the launcher: kernels after the first are chained (PDL) */
    /* inputs */
    const DynHead*     dyn;
    const LambdaPoint* pts;
    const float*    pos3;
    const float4*   par4;
    const float4*   typetab;
    const unsigned int* trips; /* [n_trips][FEP_TRIP_WORDS] */
    /* intermediates */
    float4* fsorted;
    float4* fshift_sorted;
    float2* ev2;
    double* cta_part;
    double* for_part;
    double* job_part; /* [n_red_jobs][4] */
    unsigned int* done_counter;
    /* epilogue inputs */
    const int*    atom_ptr;
    /* the atoms that receive contributions FROM THIS CONTEXT'S LIST as records {atom, first contribution, one past the
     * last, 0}, ascending by atom: one coalesced 16-byte load tells an epilogue thread all it needs (fetched before it
     * waits for the pair kernels).  heavy: more than FEP_HEAVY_MIN contributions, light: the others.  Atoms of the
     * compact numbering without any contribution (a rank's shard of a split list) are never visited: their words of
     * the result block stay zero. */
    const int4*   heavy_atoms;
    int           n_heavy;
    const int4*   light_atoms;
    int           n_light;
    const RedJob* red_jobs;
    const int*    key_job_ptr; /* [45 + G + 1]: jobs of each key, shift keys first */
    /* outputs */
    float*  res_f32;
    double* res_f64;
    PeerExchange px;
    PushTargets  push;
    unsigned long long* trace; /* NULL, or FEP_TRACE_BLOCKS x 4 global-timer stamps of the epilogue's blocks */
    unsigned int*       fault; /* host-mapped FEP_FAULT_WORDS words: why a kernel of this context trapped (fep_fault) */
};

/* A kernel that gives up (a peer that never announces its step, a bulk copy that never lands) says why in
 * host-mapped memory before it traps, so that the error the host reports names the cause instead of
 * "unspecified launch failure".  Word 0: code; 1..3: details (see fail() in fepb200_api.cu). */
#define FEP_FAULT_WORDS 4
enum { FEP_FAULT_NONE = 0, FEP_FAULT_PEER_TIMEOUT = 1, FEP_FAULT_STAGE_TIMEOUT = 2 };

/* what one step has to produce */
struct StepFlags
{
    int force, shift, energy, foreign;
};

/* Device buffers of the list builder (fep_list_build.cu); ints unless noted.  Scratch: pj, pn [P]; deg [nT+1];
 * keys, keys_out [P] of 4- or 8-byte keys; vals, vals_out, gmark, gstart [P]; th, tsc [P+1]; akeys, akeys_out,
 * avals, avals_out [P+NT]; tshift, tfirst, kshift, kgid [NT]; key_ptr [47 + G + 2].  Results: ent4 [E];
 * trips [NT][FEP_TRIP_WORDS]; tgid [NT]; orig [32 NT]; atom_ptr [nT+1] (atom_ptr[nT] = P + NS). */
struct ListBuild
{
    int *  pj, *pn, *deg, *vals, *vals_out, *gmark, *gstart, *th, *tsc, *akeys, *akeys_out, *avals, *avals_out, *tshift, *key_ptr;
    int *  tfirst, *kshift, *kgid;
    void * keys, *keys_out, *tmp;
    size_t tmp_bytes;
    int4 *        ent4;
    unsigned int* trips;
    int *         tgid, *orig, *atom_ptr;
};

#ifdef __cplusplus
extern "C" {
#endif
/* implemented in fep_kernels.cu; all launches go to `stream`; return a cudaError_t as int */
/* `events`, when not NULL, are 4 events recorded before the pass kernel and after the pass,
 * foreign and epilogue kernels (profiling mode only).  host_cur / host_pts: host copies of the
 * lambda points; beutler_mode >= 0 selects the fused Beutler kernels of fep_beutler.cu
 * (0: alphaCoul == 0; 1: one soft-core radius; 2: separate radii).  side_stream / fork_ev / join_ev
 * (may be NULL): the independent pass and foreign kernels of a large list run concurrently. */
int fep_launch_step(const KernelArgs* ka, int softcore, int elec_ewald, StepFlags sf, cudaStream_t stream,
                    long long* launch_counter, cudaEvent_t* events, const LambdaPoint* host_cur,
                    const LambdaPoint* host_pts, int beutler_mode, cudaStream_t side_stream, cudaEvent_t fork_ev,
                    cudaEvent_t join_ev);
#define FEP_FB_CTA 128
#define FEP_RING_DEPTH 4 /* trips of a warp's ring in shared memory: one evaluated, one fetched from, two on their way */
#define FEP_FB_MAXC 24
int fep_beutler_chunk_size(int n_points, int n_chunks_wanted);
int fep_beutler_ctas_per_sm(int elec_ewald, int mode, int chunk_points, int force);
/* chained != 0: a kernel of this step is already queued on `stream` and the launches may start
 * before it has completed (programmatic dependent launch, see fep_launch_kernel below) */
/* fep_gapsys.cu: the energy-only foreign-lambda passes of the Gapsys soft-core with the lambda-independent part hoisted */
int fep_gapsys_chunk_size(int n_points);
int fep_gapsys_ctas_per_sm(int elec_ewald, int chunk_points);
int fep_launch_gapsys_foreign(const KernelArgs* ka, int elec_ewald, const LambdaPoint* host_pts, cudaStream_t stream,
                              long long* launch_counter, int chained);
int fep_launch_beutler(const KernelArgs* ka, int elec_ewald, int mode, const LambdaPoint* host_cur,
                       const LambdaPoint* host_pts, int do_force, int do_foreign, int want_shift, cudaStream_t stream,
                       long long* launch_counter, int chained);
/* Sum over ranks of result blocks that live in peer-mapped memory (NVLink): out = sum_r peer[r],
 * in rank order (deterministic).  n16 = number of 16-byte words that hold fp32 data, n64 = number
 * of doubles; each peer block is [n64 doubles padded to 16 B | fp32 words]. */
#define FEP_MAX_PEERS 16
struct PeerPtrs
{
    const void* p[FEP_MAX_PEERS];
};
/* flags (may be NULL): per rank a peer-mapped array of FEP_MAX_PEERS uint32 sequence numbers; the
 * kernel then first announces `seq` in every peer's array and waits until all peers announced it
 * (the cross-GPU barrier is part of the reduction kernel). */
int fep_launch_peer_reduce(const PeerPtrs* peers, const PeerPtrs* flags, int rank, unsigned int seq, int nranks,
                           double* out_f64, int n64, size_t f64_bytes, float* out_f32, long long n32,
                           cudaStream_t stream, long long* launch_counter, int chained, unsigned int* fault);
/* the reduce-scatter flavour: this rank keeps the sum of the fp32 words [w0, w1) (its atoms' forces), of the 135
 * shift-force words at off_fshift and of the n64 doubles */
int fep_launch_peer_reduce_scatter(const PeerPtrs* peers, const PeerPtrs* flags, int rank, unsigned int seq, int nranks,
                                   double* out_f64, int n64, size_t f64_bytes, float* out_f32, long long w0, long long w1,
                                   long long off_fshift, cudaStream_t stream, long long* launch_counter, int chained,
                                   unsigned int* fault);
int fep_launch_gather_x(const float* d_x, int stride, const int* d_touched, float* pos3, int n_touched,
                        cudaStream_t stream, long long* launch_counter);
/* fepb200_export_scalars_device(): where the values sit in the fp64 part of the result block, and the
 * caller's float device buffers they are added into (any of them may be NULL) */
struct ExportLayout
{
    int ngrp, nforeign, energy, foreign, shift, atomic;
    int off_vc, off_vv, off_dvdl, off_foreign_e, off_foreign_dvdl;
};
struct ExportTargets
{
    float *eLJ, *eElec, *dvdlLJ, *dvdlElec, *eLJForeign, *eElecForeign, *dvdlLJForeign, *dvdlElecForeign, *fShift;
};
int fep_launch_export_scalars(const double* r64, const float* r32_fshift, const struct ExportLayout* lay,
                              const struct ExportTargets* targets, cudaStream_t stream, long long* launch_counter);
/* how fep_launch_add_forces() combines the result block with the caller's force array */
enum { FEP_ADD_PLAIN = 0, FEP_ADD_OVERWRITE = 1, FEP_ADD_ATOMIC = 2 };
int fep_launch_add_forces(const float* res_f32, const int* d_touched, float* d_f, int k0, int k1, int mode,
                          cudaStream_t stream, long long* launch_counter);
#ifdef __cplusplus
}
#endif

#ifdef __CUDACC__
/* Programmatic dependent launch (sm_90+).  The kernels of one step are queued back to back on one
 * stream.  Every kernel lets its successor start as soon as all of its own CTAs are running
 * (fep_pdl_launch_dependents at the top), so independent kernels (pass and foreign passes) share
 * the SMs and a dependent kernel (epilogue) has its launch latency and its prologue hidden behind
 * its predecessor's tail.  A kernel that needs its predecessors' results calls fep_pdl_wait()
 * first: it returns when the preceding kernel has completed and its writes are visible.  A kernel
 * that does NOT depend on its predecessor still calls fep_pdl_wait() as its last action, so that
 * "my predecessor in the stream has completed" implies "everything before it has completed".
 * Both instructions are no-ops in a kernel launched without the attribute. */
__device__ __forceinline__ void fep_pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void fep_pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ unsigned long long fep_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/* records the reason (first writer wins) where the host can still read it after the trap, then traps */
static __device__ __noinline__ void fep_fault(unsigned int* fault, unsigned int code, unsigned int a, unsigned int b, unsigned int c)
{
    if (fault && atomicCAS_system(fault, 0u, code) == 0u)
    {
        fault[1] = a;
        fault[2] = b;
        fault[3] = c;
        __threadfence_system();
    }
    __trap();
}

/* Shared-memory carve-out of every kernel of a step.  The driver picks a carve-out per kernel from its shared-memory
 * needs, and an SM can only change it when it is empty: a kernel whose choice differs from that of the CTAs resident
 * on an SM waits for them to drain.  That costs the overlap of a chained (PDL) kernel with its predecessor's tail,
 * and it deadlocks ranks that share one device in the tests (their epilogues spin in the cross-rank barrier while
 * another rank's pair kernel waits for the SMs to drain).  All kernels of a step therefore ask for the same one:
 * FEPB200_CARVEOUT = percent of the L1/shared array used as shared memory (default: the maximum, which the staged
 * tiles of the pair kernels want anyway); -1 = leave the choice to the driver. */
static inline int fep_carveout_percent()
{
    static const int v = [] {
        const char* e = std::getenv("FEPB200_CARVEOUT");
        return e ? std::atoi(e) : 100;
    }();
    return v;
}
/* once per kernel and device (the attribute belongs to the device's copy of the function) */
static inline void fep_prefer_carveout(const void* kernel)
{
    const int v = fep_carveout_percent();
    if (v < 0)
    {
        return;
    }
    static std::mutex                                 mtx;
    static std::set<std::pair<const void*, int>>      done;
    int                                               dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mtx);
    if (done.insert(std::make_pair(kernel, dev)).second)
    {
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v > 100 ? 100 : v);
    }
}

template<typename... KArgs, typename... Args>
static inline cudaError_t fep_launch_kernel_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem_bytes,
                                                 cudaStream_t stream, bool chained, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = grid;
    cfg.blockDim           = block;
    cfg.dynamicSmemBytes   = smem_bytes;
    cfg.stream             = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = chained ? 1 : 0;
    fep_prefer_carveout(reinterpret_cast<const void*>(kernel));
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template<typename... KArgs, typename... Args>
static inline cudaError_t fep_launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream,
                                            bool chained, Args&&... args)
{
    return fep_launch_kernel_smem(kernel, grid, block, 0, stream, chained, static_cast<Args&&>(args)...);
}
#endif

#endif
