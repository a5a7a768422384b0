/*
 * fep_kernels.cu -- the sm_100a kernels of the perturbed-pair path.
 *
 *   fep_pass_kernel     one thread per pair of the flat pair space; the pass at the current lambda
 *                       (forces + energies + dV/dlambda).  Replaces one call of
 *                       gmx_nb_free_energy_kernel() (nb_free_energy.cpp:1367) over all OpenMP
 *                       threads' lists (freeenergydispatch.cpp:189-231).
 *   fep_foreign_kernel  grid = (pair tiles) x (chunks of lambda points); evaluates the energy-only
 *                       passes of freeenergydispatch.cpp:236-306 for ALL lambda points with one
 *                       load of each pair per chunk.
 *   fep_epilogue_kernel atomic-free, deterministic: sums each atom's contiguous range of the
 *                       atom-sorted contribution buffer into per-atom forces, the shift-sorted
 *                       segment forces into shift forces, the group-sorted segment energies into
 *                       energy-group pairs, and the per-CTA partial sums into dV/dlambda and
 *                       foreign energies.  Replaces ThreadedForceBuffer::reduce
 *                       (threaded_force_buffer.cpp:320-402) and the sums at
 *                       freeenergydispatch.cpp:298-305.
 *
 * No tensor cores: the work is pairwise FP32 FMA + MUFU (rcp/rsq/lg2/ex2).  No global atomics in
 * any data path; the only atomic is the completion ticket of the epilogue.
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fep_front.cuh"

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        v += __shfl_xor_sync(FULL_MASK, v, o);
    }
    return v;
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        v += __shfl_xor_sync(FULL_MASK, v, o);
    }
    return v;
}

/* ------------------------------------------------------------------------------------------- */
/* pass at the current lambda                                                                  */
/* ------------------------------------------------------------------------------------------- */
template<int SC, bool EWALD, bool FORCE>
__global__ void __launch_bounds__(FEP_CTA) fep_pass_kernel(const __grid_constant__ KernelArgs ka, const int want_shift)
{
    __shared__ LambdaPoint s_lp;
    __shared__ float       s_red[FEP_CTA / 32][4];

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    fep_pdl_launch_dependents();
    if (tid < (int)(sizeof(LambdaPoint) / 4))
    {
        reinterpret_cast<int*>(&s_lp)[tid] = reinterpret_cast<const int*>(&ka.dyn->cur)[tid];
    }
    __syncthreads();

    /* one warp per run of trips (fep_types.h): the owner's sums are kept per lane over a segment */
    const int  t0  = ka.trip_begin + (blockIdx.x * (FEP_CTA / 32) + warp) * ka.run_trips;
    const int  t1  = min(t0 + ka.run_trips, ka.trip_end);
    const bool per_segment_energy = ka.n_gid > 1;

    float      dc = 0.0f, dv = 0.0f, vc1 = 0.0f, vv1 = 0.0f;
    FepSegment seg;
    fep_segment_clear(seg);
    for (int t = t0; t < t1; t++)
    {
        const unsigned int* tb = ka.trips + (size_t)t * FEP_TRIP_WORDS;
        const FepFetch      ft = fep_fetch<false>(ka, tb, lane);
        const FepSlot       p  = fep_slot<false>(ka, tb, ft, lane);
        float               fx = 0.0f, fy = 0.0f, fz = 0.0f, vc = 0.0f, vv = 0.0f;
        FepPair             pr;
        if (fep_fill_pair<SC>(ka, p, pr))
        {
            float fscal = 0.0f;
            if (pr.included_within)
            {
                fep_included_terms<SC, EWALD, FORCE>(ka, s_lp, pr, vc, vv, fscal, dc, dv);
            }
            float xc, fc, xv, fv;
            fep_corrections<EWALD, FORCE>(ka, pr, p.excluded, p.self, xc, fc, xv, fv);
            const float cA = pr.qq[0] * xc, cB = pr.qq[1] * xc;
            const float gA = pr.c6g[0] * xv, gB = pr.c6g[1] * xv;
            vc += s_lp.lfac_c[0] * cA + s_lp.lfac_c[1] * cB;
            vv += s_lp.lfac_v[0] * gA + s_lp.lfac_v[1] * gB;
            dc += cB - cA;
            dv += gB - gA;
            if (FORCE)
            {
                fscal += (s_lp.lfac_c[0] * pr.qq[0] + s_lp.lfac_c[1] * pr.qq[1]) * fc;
                fscal += (s_lp.lfac_v[0] * pr.c6g[0] + s_lp.lfac_v[1] * pr.c6g[1]) * fv;
                fx = fscal * p.dx;
                fy = fscal * p.dy;
                fz = fscal * p.dz;
            }
        }
        if (FORCE && p.active)
        {
            /* the partner receives -f: scattered to this pair's own slot in the atom-sorted buffer
             * (unique destination, no atomics; skipped pairs write their zero) */
            ka.fsorted[__ldg(tb + FEP_TW_DST + lane)] = make_float4(-fx, -fy, -fz, 0.0f);
        }
        /* the owner receives the sum over its segment; the segment's Vc/Vv go to its energy-group pair */
        seg.fx -= fx;
        seg.fy -= fy;
        seg.fz -= fz;
        if (per_segment_energy)
        {
            seg.vc += vc;
            seg.vv += vv;
        }
        else
        {
            vc1 += vc;
            vv1 += vv;
        }
        if (__ldg(tb + FEP_TH_FLAGS) & FEP_TRIP_LAST)
        {
            if (FORCE || per_segment_energy)
            {
                /* energy-only passes have no force slots to fill */
                if (FORCE)
                {
                    fep_segment_flush<false>(ka, tb, ft.head, seg, want_shift != 0, per_segment_energy, lane);
                }
                else
                {
                    const float svc = fep_warp_sum(seg.vc), svv = fep_warp_sum(seg.vv);
                    if (lane == 0)
                    {
                        ka.ev2[__ldg(tb + FEP_TH_SLOT_EV)] = make_float2(svc, svv);
                    }
                    fep_segment_clear(seg);
                }
            }
        }
    }

    /* dV/dlambda of this CTA (the reference accumulates one scalar per call, :1170-1178); with a single
     * energy-group pair also Vc and Vv */
    {
        const float wc = warp_sum(dc), wv = warp_sum(dv), w2 = warp_sum(vc1), w3 = warp_sum(vv1);
        if (lane == 0)
        {
            s_red[warp][0] = wc;
            s_red[warp][1] = wv;
            s_red[warp][2] = w2;
            s_red[warp][3] = w3;
        }
    }
    __syncthreads();
    if (tid < 4)
    {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < FEP_CTA / 32; w++)
        {
            s += (double)s_red[w][tid];
        }
        ka.cta_part[(size_t)tid * ka.n_cta + blockIdx.x] = s;
    }
    fep_pdl_wait(); /* first kernel of a step: no-op */
}

/* ------------------------------------------------------------------------------------------- */
/* energy-only passes at all lambda points                                                     */
/* ------------------------------------------------------------------------------------------- */

/* Sums 8 per-lane values over the 32 lanes with 9 shuffles: after three halving exchanges lane l
 * holds value index 4*(l&1) + 2*((l>>1)&1) + ((l>>2)&1), summed over all lanes. */
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane)
{
    float a[4], b[2], c;
    {
        const bool up = lane & 1;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const float send = up ? v[i] : v[i + 4];
            const float keep = up ? v[i + 4] : v[i];
            a[i]             = keep + __shfl_xor_sync(FULL_MASK, send, 1);
        }
    }
    {
        const bool up = lane & 2;
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            const float send = up ? a[i] : a[i + 2];
            const float keep = up ? a[i + 2] : a[i];
            b[i]             = keep + __shfl_xor_sync(FULL_MASK, send, 2);
        }
    }
    {
        const bool  up   = lane & 4;
        const float send = up ? b[0] : b[1];
        const float keep = up ? b[1] : b[0];
        c                = keep + __shfl_xor_sync(FULL_MASK, send, 4);
    }
    c += __shfl_xor_sync(FULL_MASK, c, 8);
    c += __shfl_xor_sync(FULL_MASK, c, 16);
    return c;
}

template<int SC, bool EWALD>
__global__ void __launch_bounds__(FEP_CTA) fep_foreign_kernel(const __grid_constant__ KernelArgs ka)
{
    __shared__ LambdaPoint s_pts[FEP_LCHUNK];
    __shared__ float       s_red[FEP_CTA / 32][3 * FEP_LCHUNK];

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int p0   = blockIdx.y * ka.chunk_points;
    const int np   = min(ka.chunk_points, ka.n_points - p0);
    fep_pdl_launch_dependents();
    if (tid < np * (int)(sizeof(LambdaPoint) / 4))
    {
        reinterpret_cast<int*>(s_pts)[tid] = reinterpret_cast<const int*>(ka.pts + p0)[tid];
    }
    __syncthreads();

    float acc_e[FEP_LCHUNK], acc_c[FEP_LCHUNK], acc_v[FEP_LCHUNK];
#pragma unroll
    for (int p = 0; p < FEP_LCHUNK; p++)
    {
        acc_e[p] = acc_c[p] = acc_v[p] = 0.0f;
    }

    /* a CTA takes a tile of trips, one warp per trip and round */
    const int t0 = ka.trip_begin + blockIdx.x * ka.tile_trips;
    const int t1 = min(t0 + ka.tile_trips, ka.trip_end);
    for (int t = t0 + warp; t < t1; t += FEP_CTA / 32)
    {
        const unsigned int* tb = ka.trips + (size_t)t * FEP_TRIP_WORDS;
        const FepFetch      ft = fep_fetch<false>(ka, tb, lane);
        const FepSlot       sl = fep_slot<false>(ka, tb, ft, lane);
        FepPair        pr;
        if (!fep_fill_pair<SC>(ka, sl, pr))
        {
            continue;
        }
        float xc, fc, xv, fv;
        fep_corrections<EWALD, false>(ka, pr, sl.excluded, sl.self, xc, fc, xv, fv);
        const float cA = pr.qq[0] * xc, cB = pr.qq[1] * xc;
        const float gA = pr.c6g[0] * xv, gB = pr.c6g[1] * xv;
        const float dcorr_c = cB - cA, dcorr_v = gB - gA;
#pragma unroll
        for (int p = 0; p < FEP_LCHUNK; p++)
        {
            if (p < np)
            {
                const LambdaPoint& lp = s_pts[p];
                float              vc = 0.0f, vv = 0.0f, fs = 0.0f, dc = dcorr_c, dv = dcorr_v;
                if (pr.included_within)
                {
                    fep_included_terms<SC, EWALD, false>(ka, lp, pr, vc, vv, fs, dc, dv);
                }
                vc += lp.lfac_c[0] * cA + lp.lfac_c[1] * cB;
                vv += lp.lfac_v[0] * gA + lp.lfac_v[1] * gB;
                acc_e[p] += vc + vv;
                acc_c[p] += dc;
                acc_v[p] += dv;
            }
        }
    }

    const float re = warp_sum8(acc_e, lane), rc = warp_sum8(acc_c, lane), rv = warp_sum8(acc_v, lane);
    if (lane < 8)
    {
        const int p         = 4 * (lane & 1) + 2 * ((lane >> 1) & 1) + ((lane >> 2) & 1);
        s_red[warp][3 * p]     = re;
        s_red[warp][3 * p + 1] = rc;
        s_red[warp][3 * p + 2] = rv;
    }
    __syncthreads();
    if (tid < 3 * np)
    {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < FEP_CTA / 32; w++)
        {
            s += (double)s_red[w][tid];
        }
        const int p = tid / 3, k = tid - 3 * p;
        ka.for_part[((size_t)(3 * (p0 + p) + k)) * ka.n_tiles + blockIdx.x] = s;
    }
    fep_pdl_wait(); /* independent of the pass kernel before it; see fep_types.h */
}

/* ------------------------------------------------------------------------------------------- */
/* epilogue                                                                                    */
/* ------------------------------------------------------------------------------------------- */
struct EpilogueLayout
{
    int atom_blocks;   /* blocks of the per-atom force sums, FEP_EPI_LANES lanes per atom    */
    int heavy_blocks;  /* blocks of the heavy atoms' force sums, one warp per atom       */
    int job_begin;     /* first reduction job handled (skips shift jobs when not asked)  */
    int job_blocks;    /* blocks for reduction jobs                                      */
    int scalar_blocks; /* blocks for dvdl (2) + Vc, Vv (2) + foreign (3*(L+1)) partial arrays */
};

__device__ __forceinline__ double block_sum_d(double v, double* s_buf)
{
    v = warp_sum_d(v);
    if ((threadIdx.x & 31) == 0)
    {
        s_buf[threadIdx.x >> 5] = v;
    }
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
    {
        for (int w = 0; w < FEP_EPI_CTA / 32; w++)
        {
            s += s_buf[w];
        }
    }
    __syncthreads();
    return s; /* valid in thread 0 */
}

/* Cross-GPU barrier of the peer exchange (fep_types.h), taken by every CTA of the epilogue right
 * after fep_pdl_wait(), i.e. when this rank's pair kernels have completed and their results are in
 * this rank's exchange slot.  Block 0 publishes them (system-scope fence, then a release store of
 * the step number `seq` into slot `rank` of every peer's flag array); every block waits until all
 * peers' announcements have arrived in the local array (acquire loads), after which the peers'
 * slots may be read.  Exactly one kernel per GPU spins and its producers have already finished, so
 * the spin cannot starve anybody; a peer that never arrives (a rank that did not launch) trips the
 * time-out and the kernel traps instead of hanging the GPU. */
/* Threads 0..nranks-1 of the calling block: (announce) release-store `seq` into slot `rank` of peer
 * threadIdx.x's flag array, then poll slot threadIdx.x of the own array until that peer has announced
 * `seq` or a later step.  All ranks must launch their steps in lockstep with the same `seq`; a peer that
 * never arrives within 4 s makes the kernel record the reason and trap (FEP_FAULT_PEER_TIMEOUT) instead
 * of hanging every GPU of the node.  Ends with a block-wide barrier. */
/* light (experiment, FEPB200_BARRIER=light; not the default): the same barrier with ONE system-scope fence on either
 * side instead of one per instruction -- the announcement is the release store alone (the data it publishes was written
 * by the preceding kernel: it is ordered before the store by the kernel boundary, and a release is cumulative), the poll
 * is a relaxed load, and one acquire fence follows the load that saw the announcement.  The idea: the barrier is 10 of
 * the 13 us the reduction kernel takes with its data already local, and system-scope fences are what it consists of.
 * Measured on 2 B200 it is SLOWER than the form with a fence before the store and acquire loads (4 us per step). */
template<typename FlagsOf>
__device__ __forceinline__ void fep_flag_barrier(FlagsOf flags_of, int rank, int nranks, unsigned int seq, bool announce,
                                                 unsigned int* fault, bool light = false)
{
    if (threadIdx.x < nranks)
    {
        if (announce)
        {
            if (!light)
            {
                __threadfence_system();
            }
            unsigned int* dst = flags_of(threadIdx.x) + rank;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(seq) : "memory");
        }
        const unsigned int* src = flags_of(rank) + threadIdx.x;
        unsigned int        v;
        unsigned long long  t0 = 0;
        for (unsigned int spins = 0;; spins++)
        {
            if (light)
            {
                asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            }
            else
            {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            }
            if ((int)(v - seq) >= 0)
            {
                if (light)
                {
                    asm volatile("fence.acq_rel.sys;" ::: "memory");
                }
                break;
            }
            if ((spins & 1023u) == 1023u)
            {
                const unsigned long long now = fep_globaltimer();
                if (t0 == 0)
                {
                    t0 = now;
                }
                else if (now - t0 > 4000000000ull) /* 4 s */
                {
                    fep_fault(fault, FEP_FAULT_PEER_TIMEOUT, (unsigned)rank, threadIdx.x, seq);
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void fep_peer_barrier(const KernelArgs& ka)
{
    fep_flag_barrier([&](int r) { return ka.px.flags[r]; }, ka.px.rank, ka.px.nranks, ka.px.seq, blockIdx.x == 0, ka.fault);
}

/* PEER: the exchange of fep_types.h.  STRONG: peer data is read with system-scope strong loads
 * (ld.volatile) instead of plain loads ordered by the barrier's acquire (FEPB200_PEER_LOAD=strong;
 * measured slower by far, kept for diagnosis). */
template<bool PEER, bool STRONG, int LANES>
__global__ void __launch_bounds__(FEP_EPI_CTA) fep_epilogue_kernel(const __grid_constant__ KernelArgs ka,
                                                                  const EpilogueLayout lay, const StepFlags sf)
{
/* PEER: element k of a sorted array lives in the exchange slot of the rank that produced it (byte
 * tables px.*_src).  The loads follow the barrier's acquire (fep_peer_barrier: ld.acquire.sys by
 * the polling threads, then bar.sync), which orders them after the producers' release. */
#define FEP_EPI_LOAD(arr, k, src) \
    (PEER ? (STRONG ? __ldcv(ka.px.arr[src] + (k)) : __ldcs(ka.px.arr[src] + (k))) : __ldcs(ka.arr + (k)))
    __shared__ double s_buf[FEP_EPI_CTA / 32];
    __shared__ bool   s_last;
    const int         tid = threadIdx.x;
    int               b   = blockIdx.x;
/* where results go: the rank's own block, or -- push reduction (PushTargets) -- the receive block of the rank that owns
 * the atom (forces) / of every rank (shift forces and scalars) */
#define FEP_EPI_FORCE_DST(atom) \
    ((!PEER && ka.push.nranks > 1) ? ka.push.f32[min((atom) / ka.push.per_rank, ka.push.nranks - 1)] : ka.res_f32)
#define FEP_EPI_STORE_ALL(part, idx, val)                 \
    do                                                    \
    {                                                     \
        if (!PEER && ka.push.nranks > 1)                  \
        {                                                 \
            for (int r_ = 0; r_ < ka.push.nranks; r_++)   \
            {                                             \
                ka.push.part[r_][idx] = (val);            \
            }                                             \
        }                                                 \
        else                                              \
        {                                                 \
            ka.res_##part[idx] = (val);                   \
        }                                                 \
    } while (0)
    /* optional trace (fepb200_epilogue_trace): per block the global timer at entry, after the wait
     * for this rank's pair kernels, after the cross-GPU barrier, and when the block's sums are done */
    unsigned long long tr0 = 0, tr1 = 0, tr2 = 0;
    if (PEER && ka.trace)
    {
        tr0 = fep_globaltimer();
    }
#define FEP_EPI_SYNC_POINT()                  \
    do                                        \
    {                                         \
        fep_pdl_wait();                       \
        if (PEER)                             \
        {                                     \
            if (ka.trace)                     \
            {                                 \
                tr1 = fep_globaltimer();      \
            }                                 \
            fep_peer_barrier(ka);             \
            if (ka.trace)                     \
            {                                 \
                tr2 = fep_globaltimer();      \
            }                                 \
        }                                     \
    } while (0)

    /* offsets into the fp64 result block: Vc[G] Vv[G] dvdl[2] foreign_E[L+1] foreign_dvdl[L+1][2] */
    const int off_vv   = ka.n_gid;
    const int off_dvdl = 2 * ka.n_gid;
    const int off_fe   = off_dvdl + 2;
    const int off_fd   = off_fe + ka.n_points;

    /* Chained behind the pair kernels (fep_types.h): what does not depend on their results -- the
     * static descriptors of this block's work -- is fetched before fep_pdl_wait(). */
    fep_pdl_launch_dependents();

    /* role order: reduction jobs and scalar sums first (few, long), per-atom gathers after, so the
     * long blocks overlap with the bulk instead of forming a tail */
    bool is_job = false;
    if (b < lay.job_blocks)
    {
        is_job           = true;
        const int    j   = lay.job_begin + b;
        const RedJob job = ka.red_jobs[j];
        /* a job is a contiguous range of at most FEP_RED_CHUNK = PER * CTA elements */
        constexpr int PER = FEP_RED_CHUNK / FEP_EPI_CTA;
        int           sr[PER]; /* producer rank of this thread's elements: static, fetched before the wait */
#pragma unroll
        for (int u = 0; u < PER; u++)
        {
            const int k = job.begin + tid + u * FEP_EPI_CTA;
            sr[u]       = (PEER && k < job.end) ? (job.kind == 0 ? ka.px.fshift_src[k] : ka.px.ev2_src[k]) : 0;
        }
        FEP_EPI_SYNC_POINT();
        double       a0 = 0.0, a1 = 0.0, a2 = 0.0;
        if (job.kind == 0)
        {
            float4 t[PER];
#pragma unroll
            for (int u = 0; u < PER; u++)
            {
                const int k = job.begin + tid + u * FEP_EPI_CTA;
                t[u]        = k < job.end ? FEP_EPI_LOAD(fshift_sorted, k, sr[u]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < PER; u++)
            {
                a0 += t[u].x;
                a1 += t[u].y;
                a2 += t[u].z;
            }
        }
        else
        {
            float2 t[PER];
#pragma unroll
            for (int u = 0; u < PER; u++)
            {
                const int k = job.begin + tid + u * FEP_EPI_CTA;
                t[u]        = k < job.end ? FEP_EPI_LOAD(ev2, k, sr[u]) : make_float2(0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < PER; u++)
            {
                a0 += t[u].x;
                a1 += t[u].y;
            }
        }
        a0 = block_sum_d(a0, s_buf);
        a1 = block_sum_d(a1, s_buf);
        a2 = block_sum_d(a2, s_buf);
        if (tid == 0)
        {
            ka.job_part[4 * (size_t)j]     = a0;
            ka.job_part[4 * (size_t)j + 1] = a1;
            ka.job_part[4 * (size_t)j + 2] = a2;
        }
    }
    else if ((b -= lay.job_blocks) < lay.scalar_blocks)
    {
        /* b = 0,1: dV/dlambda coul, vdw of the current-lambda pass; b = 2,3: its Vc, Vv (used with a single
         * energy-group pair); b = 4 + 3p + k: point p */
        const double* src;
        int           n;
        FEP_EPI_SYNC_POINT();
        if (b < 4)
        {
            src = ka.cta_part + (size_t)b * ka.n_parts;
            n   = ka.n_parts;
        }
        else
        {
            src = ka.for_part + (size_t)(b - 4) * ka.n_tiles;
            n   = ka.n_tiles;
        }
        double a = 0.0;
        if (PEER)
        {
            /* row b of every rank's partial array, ranks in order (all ranks launch the same grids) */
            const size_t row = b < 4 ? (size_t)b * ka.n_parts : (size_t)(b - 4) * ka.n_tiles;
            /* four independent (possibly remote) loads in flight per thread, summed in index order */
            const int total = n * ka.px.nranks;
            for (int k = tid; k < total; k += 4 * FEP_EPI_CTA)
            {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    const int kk = k + u * FEP_EPI_CTA;
                    v[u]         = 0.0;
                    if (kk < total)
                    {
                        const int     r = kk / n;
                        const double* q = (b < 4 ? ka.px.cta_part[r] : ka.px.for_part[r]) + row + (kk - r * n);
                        v[u]            = STRONG ? __ldcv(q) : __ldcs(q);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    a += v[u];
                }
            }
        }
        else
        {
            for (int k = tid; k < n; k += FEP_EPI_CTA)
            {
                a += src[k];
            }
        }
        a = block_sum_d(a, s_buf);
        if (tid == 0)
        {
            if (b < 2)
            {
                FEP_EPI_STORE_ALL(f64, off_dvdl + b, a);
            }
            else if (b < 4)
            {
                if (sf.energy && ka.n_gid == 1)
                {
                    FEP_EPI_STORE_ALL(f64, b == 2 ? 0 : off_vv, a);
                }
            }
            else
            {
                const int p = (b - 4) / 3, k = (b - 4) - 3 * p;
                if (k == 0)
                {
                    FEP_EPI_STORE_ALL(f64, off_fe + p, a);
                }
                else
                {
                    FEP_EPI_STORE_ALL(f64, off_fd + 2 * p + (k - 1), a);
                }
            }
        }
    }
    else if ((b -= lay.scalar_blocks) < lay.heavy_blocks)
    {
        /* atoms with long ranges (the perturbed atoms themselves: hundreds of contributions): one
         * warp per atom, FEP_HEAVY_LOADS independent loads per lane and trip, so that even over
         * NVLink the range costs one or two memory round trips instead of one per 32 contributions */
        const int lane = tid & 31;
        const int hi   = (PEER ? ka.px.heavy_begin : 0) + b * (FEP_EPI_CTA / 32) + (tid >> 5);
        const int hend = PEER ? ka.px.heavy_end : ka.n_heavy;
        int       atom = 0, k0 = 0, k1 = 0;
        int       sr[FEP_HEAVY_LOADS];
#pragma unroll
        for (int u = 0; u < FEP_HEAVY_LOADS; u++)
        {
            sr[u] = 0;
        }
        if (hi < hend)
        {
            const int4 rec = __ldg(ka.heavy_atoms + hi);
            atom = rec.x, k0 = rec.y, k1 = rec.z;
            if (PEER)
            {
#pragma unroll
                for (int u = 0; u < FEP_HEAVY_LOADS; u++)
                {
                    const int k = k0 + lane + 32 * u;
                    sr[u]       = k < k1 ? ka.px.slot_src[k] : 0;
                }
            }
        }
        FEP_EPI_SYNC_POINT();
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;
        for (int k = k0 + lane; k < k1; k += 32 * FEP_HEAVY_LOADS)
        {
            float4 t[FEP_HEAVY_LOADS];
            if (PEER && k != k0 + lane)
            {
#pragma unroll
                for (int u = 0; u < FEP_HEAVY_LOADS; u++)
                {
                    sr[u] = (k + 32 * u < k1) ? ka.px.slot_src[k + 32 * u] : 0;
                }
            }
#pragma unroll
            for (int u = 0; u < FEP_HEAVY_LOADS; u++)
            {
                t[u] = (k + 32 * u < k1) ? FEP_EPI_LOAD(fsorted, k + 32 * u, sr[u]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < FEP_HEAVY_LOADS; u++)
            {
                fx += t[u].x;
                fy += t[u].y;
                fz += t[u].z;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            fx += __shfl_xor_sync(FULL_MASK, fx, o);
            fy += __shfl_xor_sync(FULL_MASK, fy, o);
            fz += __shfl_xor_sync(FULL_MASK, fz, o);
        }
        if (hi < hend && lane < 3)
        {
            FEP_EPI_FORCE_DST(atom)[3 * (size_t)atom + lane] = lane == 0 ? fx : (lane == 1 ? fy : fz);
        }
    }
    else
    {
        /* FEP_EPI_LANES lanes per touched atom; contributions are visited in ascending index order by
         * lane stride, then combined with a fixed xor tree: deterministic */
        b -= lay.heavy_blocks;
        /* PEER: this rank sums the atoms it owns (the forces are reduce-scattered by atom range);
         * a remote round trip costs the same for 4 lanes as for 8, and half the threads means half
         * the waves of blocks */
        const int li   = (PEER ? ka.px.light_begin : 0) + b * (FEP_EPI_CTA / LANES) + (tid / LANES);
        const int lend = PEER ? ka.px.light_end : ka.n_light;
        const int sub  = tid % LANES;
        float     fx = 0.0f, fy = 0.0f, fz = 0.0f;
        int       atom = 0, k0 = 0, k1 = 0;
        int       sr[4] = { 0, 0, 0, 0 }; /* producer ranks of the first trip's elements (static) */
        const bool mine = li < lend;
        if (mine)
        {
            const int4 rec = __ldg(ka.light_atoms + li);
            atom = rec.x, k0 = rec.y, k1 = rec.z;
            if (PEER)
            {
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    const int k = k0 + sub + LANES * u;
                    sr[u]       = k < k1 ? ka.px.slot_src[k] : 0;
                }
            }
        }
        FEP_EPI_SYNC_POINT();
        if (mine)
        {
            /* the atom's contributions are contiguous in fsorted: the lanes stream them, four
             * independent 16-byte loads per lane and trip (most atoms need a single trip) */
            for (int k = k0 + sub; k < k1; k += 4 * LANES)
            {
                float4 t[4];
                if (PEER && k != k0 + sub)
                {
#pragma unroll
                    for (int u = 0; u < 4; u++)
                    {
                        sr[u] = (k + LANES * u < k1) ? ka.px.slot_src[k + LANES * u] : 0;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    t[u] = (k + LANES * u < k1) ? FEP_EPI_LOAD(fsorted, k + LANES * u, sr[u])
                                                : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    fx += t[u].x;
                    fy += t[u].y;
                    fz += t[u].z;
                }
            }
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1)
        {
            fx += __shfl_xor_sync(FULL_MASK, fx, o);
            fy += __shfl_xor_sync(FULL_MASK, fy, o);
            fz += __shfl_xor_sync(FULL_MASK, fz, o);
        }
        if (mine && sub < 3)
        {
            FEP_EPI_FORCE_DST(atom)[3 * (size_t)atom + sub] = sub == 0 ? fx : (sub == 1 ? fy : fz);
        }
    }

    if (PEER && ka.trace && tid == 0 && blockIdx.x < FEP_TRACE_BLOCKS)
    {
        /* role of the block in the low two bits of the first stamp: 0 job, 1 scalar, 2 heavy atoms, 3 atoms */
        const int role = (int)blockIdx.x < lay.job_blocks                                        ? 0
                         : (int)blockIdx.x < lay.job_blocks + lay.scalar_blocks                  ? 1
                         : (int)blockIdx.x < lay.job_blocks + lay.scalar_blocks + lay.heavy_blocks ? 2
                                                                                                  : 3;
        unsigned long long* t = ka.trace + 4 * (size_t)blockIdx.x;
        t[0] = (tr0 & ~3ull) | (unsigned long long)role;
        t[1] = tr1;
        t[2] = tr2;
        t[3] = fep_globaltimer();
    }
    /* the job block that finishes last adds up the job partials per output key, in job order;
     * only job blocks take a ticket (bar.sync + one fence by the ticket thread orders the block's
     * partial before the ticket) */
    if (lay.job_blocks == 0)
    {
        /* nothing to add up (empty shard): the first block writes the zeros */
        if (blockIdx.x != 0)
        {
            return;
        }
    }
    else
    {
        if (!is_job)
        {
            return;
        }
        __syncthreads();
        if (tid == 0)
        {
            __threadfence();
            const unsigned ticket = atomicAdd(ka.done_counter, 1u);
            s_last                = (ticket == (unsigned)lay.job_blocks - 1u);
        }
        __syncthreads();
        if (!s_last)
        {
            return;
        }
        __threadfence();
    }
    /* one thread per output value, four loads of job partials in flight at a time (a key can have dozens of
     * jobs -- the central shift vector, a populated energy-group pair -- and a serial walk over them by a few
     * threads was the tail of the whole step); energy-group sums only when there is more than one pair */
    {
        const int n_sh = sf.shift ? 3 * FEP_NUM_SHIFT : 0;
        const int n_en = (sf.energy && ka.n_gid > 1) ? 2 * ka.n_gid : 0;
        for (int o = tid; o < n_sh + n_en; o += FEP_EPI_CTA)
        {
            int key, comp;
            if (o < n_sh)
            {
                key  = o / 3;
                comp = o - 3 * key;
            }
            else
            {
                const int e = o - n_sh;
                comp        = e >= ka.n_gid ? 1 : 0;
                key         = FEP_NUM_SHIFT + (e - comp * ka.n_gid);
            }
            const int j1 = ka.key_job_ptr[key + 1];
            double    a  = 0.0;
            for (int j = ka.key_job_ptr[key]; j < j1; j += 4)
            {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    v[u] = (j + u < j1) ? __ldcg(ka.job_part + 4 * (size_t)(j + u) + comp) : 0.0;
                }
                a += (v[0] + v[1]) + (v[2] + v[3]);
            }
            if (o < n_sh)
            {
                FEP_EPI_STORE_ALL(f32, 3 * (size_t)ka.n_touched + o, (float)a);
            }
            else
            {
                FEP_EPI_STORE_ALL(f64, (comp ? off_vv : 0) + (key - FEP_NUM_SHIFT), a);
            }
        }
    }
    if (tid == 0 && lay.job_blocks != 0)
    {
        *ka.done_counter = 0u;
    }
#undef FEP_EPI_LOAD
#undef FEP_EPI_SYNC_POINT
#undef FEP_EPI_FORCE_DST
#undef FEP_EPI_STORE_ALL
}

/* coordinates of the touched atoms from a device-resident array of `stride` floats per atom:
 * rvec[natoms] (stride 3) or the xyzq float4[natoms] of the nbnxm GPU atom data (stride 4) */
__global__ void __launch_bounds__(256) fep_gather_x_kernel(const float* __restrict__ x, int stride,
                                                          const int* __restrict__ touched, float* __restrict__ pos3, int n)
{
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k < n)
    {
        const size_t a = (size_t)stride * (size_t)touched[k];
        pos3[3 * (size_t)k]     = x[a];
        pos3[3 * (size_t)k + 1] = x[a + 1];
        pos3[3 * (size_t)k + 2] = x[a + 2];
    }
}

/* forces of the touched atoms [k0, k1) of the result block added into (or stored to) a
 * device-resident rvec[natoms] array; every atom occurs once, so there is nothing to synchronise */
__global__ void __launch_bounds__(256) fep_add_forces_kernel(const float* __restrict__ res_f32, const int* __restrict__ touched,
                                                            float* __restrict__ f, int k0, int k1, int mode)
{
    const int k = k0 + blockIdx.x * 256 + threadIdx.x;
    if (k < k1)
    {
        const size_t a  = (size_t)touched[k];
        const float  fx = res_f32[3 * (size_t)k], fy = res_f32[3 * (size_t)k + 1], fz = res_f32[3 * (size_t)k + 2];
        if (mode == FEP_ADD_ATOMIC) /* another stream may be adding into f at the same time (FEPB200_ATOMIC_OUTPUTS) */
        {
            atomicAdd(f + 3 * a, fx);
            atomicAdd(f + 3 * a + 1, fy);
            atomicAdd(f + 3 * a + 2, fz);
        }
        else if (mode == FEP_ADD_OVERWRITE)
        {
            f[3 * a]     = fx;
            f[3 * a + 1] = fy;
            f[3 * a + 2] = fz;
        }
        else
        {
            f[3 * a] += fx;
            f[3 * a + 1] += fy;
            f[3 * a + 2] += fz;
        }
    }
}

/* the scalars and shift forces of the result block added into the float device buffers the fork's nbnxm
 * GPU module reduces from (NBAtomDataGpu::eLJ, eElec, dvdlLJ, dvdlElec, e*Foreign[L+1], dvdl*Foreign[L+1],
 * fShift[45]); one block, one writer per element.  Within one stream there is nothing to synchronise; with
 * lay.atomic (FEPB200_ATOMIC_OUTPUTS) the additions are atomic, because the fork's kernels of the other
 * locality add into the same buffers from their own stream (atomicAdd, nbnxm_cuda_kernel_utils.cuh) */
__device__ __forceinline__ void fep_export_add(float* target, float v, int atomic)
{
    if (atomic)
    {
        atomicAdd(target, v);
    }
    else
    {
        *target += v;
    }
}

__global__ void __launch_bounds__(256) fep_export_scalars_kernel(const double* __restrict__ r64, const float* __restrict__ r32_fshift,
                                                                ExportLayout lay, ExportTargets t)
{
    const int tid = threadIdx.x;
    if (tid == 0 && lay.energy)
    {
        double vc = 0.0, vv = 0.0;
        for (int g = 0; g < lay.ngrp; g++)
        {
            vc += r64[lay.off_vc + g];
            vv += r64[lay.off_vv + g];
        }
        if (t.eElec) fep_export_add(t.eElec, (float)vc, lay.atomic);
        if (t.eLJ) fep_export_add(t.eLJ, (float)vv, lay.atomic);
    }
    if (tid == 1)
    {
        if (t.dvdlElec) fep_export_add(t.dvdlElec, (float)r64[lay.off_dvdl], lay.atomic);
        if (t.dvdlLJ) fep_export_add(t.dvdlLJ, (float)r64[lay.off_dvdl + 1], lay.atomic);
    }
    if (lay.foreign)
    {
        /* the fork keeps a Coulomb and an LJ share per lambda point and adds both into the same foreign term
         * (gpu_common.h:176-191); the energy of a point is exported whole through the LJ share */
        for (int i = tid; i <= lay.nforeign; i += blockDim.x)
        {
            if (t.eLJForeign) fep_export_add(t.eLJForeign + i, (float)r64[lay.off_foreign_e + i], lay.atomic);
            if (t.dvdlElecForeign) fep_export_add(t.dvdlElecForeign + i, (float)r64[lay.off_foreign_dvdl + 2 * i], lay.atomic);
            if (t.dvdlLJForeign) fep_export_add(t.dvdlLJForeign + i, (float)r64[lay.off_foreign_dvdl + 2 * i + 1], lay.atomic);
        }
    }
    if (lay.shift && t.fShift)
    {
        for (int i = tid; i < 3 * FEP_NUM_SHIFT; i += blockDim.x)
        {
            fep_export_add(t.fShift + i, r32_fshift[i], lay.atomic);
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* launchers                                                                                   */
/* ------------------------------------------------------------------------------------------- */
/* FEPB200_OVERLAP = pdl (default) | streams | none */
static int fep_overlap_mode()
{
    static const int mode = [] {
        const char* e = std::getenv("FEPB200_OVERLAP");
        if (e && std::strcmp(e, "streams") == 0)
        {
            return 1;
        }
        if (e && std::strcmp(e, "none") == 0)
        {
            return 0;
        }
        return 2;
    }();
    return mode;
}

template<int SC, bool EWALD>
static cudaError_t launch_variants(KernelArgs& ka, StepFlags sf, cudaStream_t stream, long long* counter,
                                   cudaEvent_t* ev, const LambdaPoint* host_cur, const LambdaPoint* host_pts,
                                   int beutler_mode, cudaStream_t side, cudaEvent_t fork_ev, cudaEvent_t join_ev,
                                   bool* chain_out)
{
    const bool foreign = sf.foreign && ka.n_points > 0;
    const bool beutler = SC == FEP_SC_BEUTLER && beutler_mode >= 0;
    /* fep_beutler.cu.  Small lists are latency-bound: one launch computes the pass at the current
     * lambda and the first chunk of foreign points from one load of each pair.  On large lists the
     * fused kernel's register count costs more occupancy than the shared load saves, so the pass
     * and the foreign chunks are separate launches of the same code. */
    const bool one_launch = beutler && sf.force && foreign && ka.fuse_pass_and_foreign;
    int        rc         = 0;
    /* How the independent pass and foreign kernels of a large list share the GPU, and how the
     * epilogue follows them: "pdl" (default) chains all kernels of the step on one stream with
     * programmatic dependent launches (fep_types.h); "streams" forks the foreign kernels to a side
     * stream; "none" queues them plainly.  Profiling mode times every kernel alone. */
    const int  mode   = ev ? 0 : fep_overlap_mode();
    const bool pdl    = mode == 2;
    bool       queued = false; /* a kernel of this step is on `stream` */
    ka.pdl_chain      = pdl ? 1 : 0;
    if (ev)
    {
        cudaEventRecord(ev[0], stream);
    }
    ka.n_parts = 0;
    const bool use_side = mode == 1 && side != nullptr && beutler && sf.force && foreign && !one_launch;
    if (use_side)
    {
        cudaEventRecord(fork_ev, stream);
    }
    if (ka.n_pairs > 0 && !one_launch)
    {
        if (beutler && sf.force)
        {
            rc = fep_launch_beutler(&ka, EWALD ? 1 : 0, beutler_mode, host_cur, host_pts, 1, 0, sf.shift, stream, counter, 0);
            ka.n_parts = ka.pass_n_tiles * (FEP_FB_CTA / 32); /* four sums per warp */
        }
        else
        {
            /* generic pass kernel, one thread per pair */
            if (sf.force)
            {
                fep_launch_kernel(fep_pass_kernel<SC, EWALD, true>, dim3(ka.n_cta), dim3(FEP_CTA), stream, false, ka, sf.shift);
            }
            else
            {
                fep_launch_kernel(fep_pass_kernel<SC, EWALD, false>, dim3(ka.n_cta), dim3(FEP_CTA), stream, false, ka, 0);
            }
            ka.n_parts = ka.n_cta;
            (*counter)++;
        }
        queued = true;
    }
    if (ev)
    {
        cudaEventRecord(ev[1], stream);
    }
    if (rc == 0 && ka.n_pairs > 0 && foreign)
    {
        if (beutler)
        {
            const bool   overlap = use_side && ka.n_parts > 0;
            cudaStream_t fstream = stream;
            if (overlap)
            {
                /* fork_ev was recorded on `stream` before the pass kernel was queued */
                cudaStreamWaitEvent(side, fork_ev, 0);
                fstream = side;
            }
            rc = fep_launch_beutler(&ka, EWALD ? 1 : 0, beutler_mode, host_cur, host_pts, one_launch ? 1 : 0, 1,
                                    one_launch ? sf.shift : 0, fstream, counter, (pdl && queued) ? 1 : 0);
            if (overlap)
            {
                cudaEventRecord(join_ev, side);
                cudaStreamWaitEvent(stream, join_ev, 0);
            }
            if (one_launch)
            {
                ka.n_parts = ka.n_tiles;
            }
        }
        else if (SC == FEP_SC_GAPSYS && ka.gapsys_hoisted)
        {
            rc = fep_launch_gapsys_foreign(&ka, EWALD ? 1 : 0, host_pts, stream, counter, (pdl && queued) ? 1 : 0);
        }
        else
        {
            const dim3 grid(ka.n_tiles, ka.n_chunks);
            fep_launch_kernel(fep_foreign_kernel<SC, EWALD>, grid, dim3(FEP_CTA), stream, pdl && queued, ka);
            (*counter)++;
        }
        queued = true;
    }
    if (rc != 0)
    {
        return (cudaError_t)rc;
    }
    if (ev)
    {
        cudaEventRecord(ev[2], stream);
    }
    *chain_out = pdl && queued;
    return cudaGetLastError();
}

extern "C" int fep_launch_step(const KernelArgs* kap, int softcore, int elec_ewald, StepFlags sf, cudaStream_t stream,
                               long long* counter, cudaEvent_t* ev, const LambdaPoint* host_cur,
                               const LambdaPoint* host_pts, int beutler_mode, cudaStream_t side, cudaEvent_t fork_ev,
                               cudaEvent_t join_ev)
{
    KernelArgs  ka = *kap; /* local copy: n_parts depends on which pass kernel ran */
    cudaError_t err;
    bool        chain = false;
    switch (softcore * 2 + (elec_ewald ? 1 : 0))
    {
#define FEP_CASE(SCV, EW) \
    case SCV * 2 + (EW ? 1 : 0): \
        err = launch_variants<SCV, EW>(ka, sf, stream, counter, ev, host_cur, host_pts, beutler_mode, side, fork_ev, \
                                        join_ev, &chain); \
        break;
        FEP_CASE(FEP_SC_NONE, false)
        FEP_CASE(FEP_SC_NONE, true)
        FEP_CASE(FEP_SC_BEUTLER, false)
        FEP_CASE(FEP_SC_BEUTLER, true)
        FEP_CASE(FEP_SC_GAPSYS, false)
        FEP_CASE(FEP_SC_GAPSYS, true)
#undef FEP_CASE
        default: return (int)cudaErrorInvalidValue;
    }
    if (err != cudaSuccess)
    {
        return (int)err;
    }
    const bool     peer    = ka.px.nranks > 1;
    const int      n_atoms = peer ? ka.px.light_end - ka.px.light_begin : ka.n_light;
    EpilogueLayout lay;
    const int      n_heavy = peer ? ka.px.heavy_end - ka.px.heavy_begin : ka.n_heavy;
    /* lanes per light atom: FEP_EPI_LANES, or FEPB200_EPI_LANES = 2 | 4 | 8 (experiment: with the heavy
     * atoms in their own role, fewer lanes mean fewer blocks and waves; the same value must be used on
     * every rank, and it fixes the summation order, i.e. the last bits of the forces) */
    static const int lanes = [] {
        const char* e = std::getenv("FEPB200_EPI_LANES");
        const int   v = e ? std::atoi(e) : 0;
        return (v == 2 || v == 4 || v == 8) ? v : FEP_EPI_LANES;
    }();
    const int      per_blk = FEP_EPI_CTA / lanes;
    lay.atom_blocks   = sf.force ? (n_atoms + per_blk - 1) / per_blk : 0;
    lay.heavy_blocks  = sf.force ? (n_heavy + FEP_EPI_CTA / 32 - 1) / (FEP_EPI_CTA / 32) : 0;
    /* jobs are ordered shift jobs first, then energy-group jobs */
    const int j0      = sf.shift ? 0 : ka.n_shift_jobs;
    const int j1      = (sf.energy && ka.n_gid > 1) ? ka.n_red_jobs : ka.n_shift_jobs; /* one pair: summed per CTA instead */
    lay.job_begin     = j0;
    lay.job_blocks    = j1 > j0 ? j1 - j0 : 0;
    lay.scalar_blocks = 4 + ((sf.foreign && ka.n_points > 0) ? 3 * ka.n_points : 0);
    const int blocks  = lay.atom_blocks + lay.heavy_blocks + lay.job_blocks + lay.scalar_blocks;
    const dim3 grid(blocks), block(FEP_EPI_CTA);
#define FEP_EPI_LAUNCH(P, S)                                                                                 \
    do                                                                                                       \
    {                                                                                                        \
        if (lanes == 2)                                                                                      \
        {                                                                                                    \
            fep_launch_kernel(fep_epilogue_kernel<P, S, 2>, grid, block, stream, chain, ka, lay, sf);        \
        }                                                                                                    \
        else if (lanes == 4)                                                                                 \
        {                                                                                                    \
            fep_launch_kernel(fep_epilogue_kernel<P, S, 4>, grid, block, stream, chain, ka, lay, sf);        \
        }                                                                                                    \
        else                                                                                                 \
        {                                                                                                    \
            fep_launch_kernel(fep_epilogue_kernel<P, S, FEP_EPI_LANES>, grid, block, stream, chain, ka, lay, sf); \
        }                                                                                                    \
    } while (0)
    if (peer)
    {
        static const bool strong = [] {
            const char* e = std::getenv("FEPB200_PEER_LOAD");
            return e && std::strcmp(e, "strong") == 0;
        }();
        if (strong)
        {
            FEP_EPI_LAUNCH(true, true);
        }
        else
        {
            FEP_EPI_LAUNCH(true, false);
        }
    }
    else
    {
        FEP_EPI_LAUNCH(false, false);
    }
#undef FEP_EPI_LAUNCH
    (*counter)++;
    if (ev)
    {
        cudaEventRecord(ev[3], stream);
    }
    return (int)cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------- */
/* multi-GPU: one-shot reduction over peer memory                                              */
/* ------------------------------------------------------------------------------------------- */
/* FEPB200_BARRIER=light: the reduction kernels' cross-GPU barrier in its light form (fep_flag_barrier; experiment,
 * measured SLOWER on 2 B200: 57.8 against 53.3 us per C5 step, profiles/r02_multi_gpu_push_and_barrier.txt) */
static int fep_light_barrier()
{
    static const int light = [] {
        const char* e = std::getenv("FEPB200_BARRIER");
        return (e && std::strcmp(e, "light") == 0) ? 1 : 0;
    }();
    return light;
}

/* Every rank has published its result block [f64 | f32] in a buffer all ranks have mapped
 * (NVLink peer access).  Each rank reads every block once and keeps the full sum: a one-shot
 * all-reduce, latency-optimal for the ~1 MB blocks of this path; the order of the additions is the
 * rank order on every rank, so all ranks obtain bit-identical results. */
template<int NR>
__global__ void __launch_bounds__(256) fep_peer_reduce_kernel(const __grid_constant__ PeerPtrs peers,
                                                             const __grid_constant__ PeerPtrs flags, int rank,
                                                             unsigned int seq, int nranks, double* __restrict__ out_f64,
                                                             int n64, size_t f64_bytes, float* __restrict__ out_f32,
                                                             long long n32, unsigned int* fault, int light_barrier)
{
    fep_pdl_wait(); /* chained behind the epilogue that completes this rank's block */
    if (flags.p[0] != nullptr)
    {
        /* Cross-GPU barrier inside the reduction kernel.  This rank's block was completed by the
         * previous kernel on this stream; announce step `seq` in every peer's flag array (slot
         * `rank`), then wait until every peer's announcement has arrived in ours.  One kernel per
         * GPU takes part, so the spin cannot starve a producer on the same device. */
        fep_flag_barrier([&](int r) { return static_cast<unsigned int*>(const_cast<void*>(flags.p[r])); }, rank, nranks, seq,
                         blockIdx.x == 0, fault, light_barrier != 0);
    }
    /* NR > 0: compile-time rank count, all NR remote loads are in flight before the first add
     * (a peer load is ~1.5 us; serialising them would cost NR times that) */
    const int       nr = NR > 0 ? NR : nranks;
    const long long i  = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long n4 = n32 >> 2;
    if (i < n4)
    {
        float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (NR > 0)
        {
            float4 v[NR > 0 ? NR : 1];
#pragma unroll
            for (int r = 0; r < NR; r++)
            {
                /* written by another GPU: never from a stale cache line */
                v[r] = __ldcv(reinterpret_cast<const float4*>(static_cast<const char*>(peers.p[r]) + f64_bytes) + i);
            }
#pragma unroll
            for (int r = 0; r < NR; r++)
            {
                a.x += v[r].x;
                a.y += v[r].y;
                a.z += v[r].z;
                a.w += v[r].w;
            }
        }
        else
        {
            for (int r = 0; r < nr; r++)
            {
                const float4 v = __ldcv(reinterpret_cast<const float4*>(static_cast<const char*>(peers.p[r]) + f64_bytes) + i);
                a.x += v.x;
                a.y += v.y;
                a.z += v.z;
                a.w += v.w;
            }
        }
        reinterpret_cast<float4*>(out_f32)[i] = a;
    }
    else if (i < n4 + (n32 & 3))
    {
        const long long j = 4 * n4 + (i - n4);
        float           a = 0.0f;
        for (int r = 0; r < nr; r++)
        {
            a += __ldcv(reinterpret_cast<const float*>(static_cast<const char*>(peers.p[r]) + f64_bytes) + j);
        }
        out_f32[j] = a;
    }
    if (i < n64)
    {
        double a = 0.0;
        for (int r = 0; r < nr; r++)
        {
            a += __ldcv(static_cast<const double*>(peers.p[r]) + i);
        }
        out_f64[i] = a;
    }
}

extern "C" int fep_launch_peer_reduce(const PeerPtrs* peers, const PeerPtrs* flagsp, int rank, unsigned int seq,
                                      int nranks, double* out_f64, int n64, size_t f64_bytes, float* out_f32,
                                      long long n32, cudaStream_t stream, long long* counter, int chained,
                                      unsigned int* fault)
{
    PeerPtrs noflags{};
    const PeerPtrs* flags = flagsp ? flagsp : &noflags;
    const long long items  = (n32 >> 2) + (n32 & 3);
    const unsigned  blocks = (unsigned)((std::max<long long>(items, n64) + 255) / 256);
    switch (nranks)
    {
        case 2: fep_launch_kernel(fep_peer_reduce_kernel<2>, dim3(blocks), dim3(256), stream, chained != 0, *peers, *flags, rank, seq, nranks, out_f64, n64, f64_bytes, out_f32, n32, fault, fep_light_barrier()); break;
        case 4: fep_launch_kernel(fep_peer_reduce_kernel<4>, dim3(blocks), dim3(256), stream, chained != 0, *peers, *flags, rank, seq, nranks, out_f64, n64, f64_bytes, out_f32, n32, fault, fep_light_barrier()); break;
        case 8: fep_launch_kernel(fep_peer_reduce_kernel<8>, dim3(blocks), dim3(256), stream, chained != 0, *peers, *flags, rank, seq, nranks, out_f64, n64, f64_bytes, out_f32, n32, fault, fep_light_barrier()); break;
        default: fep_launch_kernel(fep_peer_reduce_kernel<0>, dim3(blocks), dim3(256), stream, chained != 0, *peers, *flags, rank, seq, nranks, out_f64, n64, f64_bytes, out_f32, n32, fault, fep_light_barrier()); break;
    }
    (*counter)++;
    return (int)cudaGetLastError();
}

/* Reduce-scatter of the ranks' partial result blocks over NVLink peer memory, with the all-reduce of everything that
 * is small: rank r sums the forces of ITS atoms (compact range [a0, a1), a0 a multiple of 4) over all ranks' blocks --
 * (N-1)/N of one block crosses NVLink per rank, where the all-reduce of fep_peer_reduce_kernel pulls N-1 whole
 * blocks -- and every rank sums the 45 shift forces and the fp64 scalars of all blocks.  Sums in rank order, so the
 * shift forces and scalars are bit-identical on all ranks.  The barrier is the one of fep_peer_reduce_kernel.
 * Items: [0, n4) float4 of the owned force range, [n4, n4 + n_tail) its last words when the range does not end on a
 * 16-byte boundary, then 135 shift-force words, then n64 doubles. */
template<int NR>
__global__ void __launch_bounds__(256) fep_peer_reduce_scatter_kernel(const __grid_constant__ PeerPtrs peers,
                                                                     const __grid_constant__ PeerPtrs flags, int rank,
                                                                     unsigned int seq, int nranks, double* __restrict__ out_f64,
                                                                     int n64, size_t f64_bytes, float* __restrict__ out_f32,
                                                                     long long w0, long long w1, long long off_fshift,
                                                                     unsigned int* fault, int light_barrier)
{
    fep_pdl_wait(); /* chained behind the epilogue that completes this rank's block */
    if (flags.p[0] != nullptr)
    {
        fep_flag_barrier([&](int r) { return static_cast<unsigned int*>(const_cast<void*>(flags.p[r])); }, rank, nranks, seq,
                         blockIdx.x == 0, fault, light_barrier != 0);
    }
    const int       nr     = NR > 0 ? NR : nranks;
    const long long n4     = (w1 - w0) >> 2;
    const long long n_tail = (w1 - w0) & 3;
    const long long i      = (long long)blockIdx.x * 256 + threadIdx.x;
    auto            f32_of = [&](int r) { return reinterpret_cast<const float*>(static_cast<const char*>(peers.p[r]) + f64_bytes); };
    if (i < n4)
    {
        float4 v[NR > 0 ? NR : FEP_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < (NR > 0 ? NR : FEP_MAX_PEERS); r++)
        {
            /* all loads in flight before the first add (a peer load is ~1.5 us); written by another GPU: never
             * from a stale cache line */
            v[r] = r < nr ? __ldcv(reinterpret_cast<const float4*>(f32_of(r) + w0) + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int r = 0; r < (NR > 0 ? NR : FEP_MAX_PEERS); r++)
        {
            a.x += v[r].x;
            a.y += v[r].y;
            a.z += v[r].z;
            a.w += v[r].w;
        }
        reinterpret_cast<float4*>(out_f32 + w0)[i] = a;
    }
    else if (i < n4 + n_tail + 3 * FEP_NUM_SHIFT)
    {
        const long long j = i - n4;
        const long long w = j < n_tail ? w0 + 4 * n4 + j : off_fshift + (j - n_tail);
        float           a = 0.0f;
        for (int r = 0; r < nr; r++)
        {
            a += __ldcv(f32_of(r) + w);
        }
        out_f32[w] = a;
    }
    else if (i < n4 + n_tail + 3 * FEP_NUM_SHIFT + n64)
    {
        const long long j = i - (n4 + n_tail + 3 * FEP_NUM_SHIFT);
        double          a = 0.0;
        for (int r = 0; r < nr; r++)
        {
            a += __ldcv(static_cast<const double*>(peers.p[r]) + j);
        }
        out_f64[j] = a;
    }
}

extern "C" int fep_launch_peer_reduce_scatter(const PeerPtrs* peers, const PeerPtrs* flagsp, int rank, unsigned int seq,
                                              int nranks, double* out_f64, int n64, size_t f64_bytes, float* out_f32,
                                              long long w0, long long w1, long long off_fshift, cudaStream_t stream,
                                              long long* counter, int chained, unsigned int* fault)
{
    PeerPtrs        noflags{};
    const PeerPtrs* flags  = flagsp ? flagsp : &noflags;
    const long long items  = ((w1 - w0) >> 2) + ((w1 - w0) & 3) + 3 * FEP_NUM_SHIFT + n64;
    const unsigned  blocks = (unsigned)((items + 255) / 256);
#define FEP_RS_LAUNCH(N)                                                                                                  \
    fep_launch_kernel(fep_peer_reduce_scatter_kernel<N>, dim3(blocks), dim3(256), stream, chained != 0, *peers, *flags, rank, \
                      seq, nranks, out_f64, n64, f64_bytes, out_f32, w0, w1, off_fshift, fault, fep_light_barrier())
    switch (nranks)
    {
        case 2: FEP_RS_LAUNCH(2); break;
        case 4: FEP_RS_LAUNCH(4); break;
        case 8: FEP_RS_LAUNCH(8); break;
        default: FEP_RS_LAUNCH(0); break;
    }
#undef FEP_RS_LAUNCH
    (*counter)++;
    return (int)cudaGetLastError();
}

extern "C" int fep_launch_add_forces(const float* res_f32, const int* d_touched, float* d_f, int k0, int k1, int mode,
                                     cudaStream_t stream, long long* counter)
{
    if (k1 > k0)
    {
        fep_add_forces_kernel<<<(k1 - k0 + 255) / 256, 256, 0, stream>>>(res_f32, d_touched, d_f, k0, k1, mode);
        (*counter)++;
    }
    return (int)cudaGetLastError();
}

extern "C" int fep_launch_export_scalars(const double* r64, const float* r32_fshift, const ExportLayout* lay,
                                         const ExportTargets* targets, cudaStream_t stream, long long* counter)
{
    fep_export_scalars_kernel<<<1, 256, 0, stream>>>(r64, r32_fshift, *lay, *targets);
    (*counter)++;
    return (int)cudaGetLastError();
}

extern "C" int fep_launch_gather_x(const float* d_x, int stride, const int* d_touched, float* pos3, int n_touched,
                                   cudaStream_t stream, long long* counter)
{
    if (n_touched > 0)
    {
        fep_gather_x_kernel<<<(n_touched + 255) / 256, 256, 0, stream>>>(d_x, stride, d_touched, pos3, n_touched);
        (*counter)++;
    }
    return (int)cudaGetLastError();
}
