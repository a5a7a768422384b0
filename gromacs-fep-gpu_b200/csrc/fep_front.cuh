/*
 * fep_front.cuh -- how the pair kernels get their input: the front end shared by fep_beutler.cu and
 * fep_kernels.cu on the trip layout of fep_types.h.
 *
 *   - A WARP evaluates whole runs of trips (fep_types.h), run after run with a grid-wide stride, one trip
 *     after the other; nothing in the loop involves the other warps of its CTA (no bar.sync).
 *   - Staging: each warp owns a ring of FEP_RING_DEPTH trip blocks in shared memory.  One lane brings the
 *     block of the trip FEP_RING_DEPTH - 1 ahead in with ONE bulk copy (cp.async.bulk, completion on the
 *     ring slot's mbarrier): the list stream costs no registers, no address arithmetic and no load
 *     instructions in the loop, and is in flight two trips before it is needed.
 *   - Per trip the warp-uniform data (owner coordinates + shift vector; owner charges in the trip's header)
 *     replace what the reference sets up per i-entry (nb_free_energy.cpp:466-503).
 *   - The dependent per-pair loads -- the partner's coordinates -- are issued one trip ahead (FepFetch);
 *     the two type-table rows come from a table that lives in L1.
 */
#ifndef FEPB200_FEP_FRONT_CUH
#define FEPB200_FEP_FRONT_CUH

#include "fep_pair_math.cuh"

#define FEP_FULL_MASK 0xffffffffu

/* ---- mbarrier + bulk copy (sm_90+; SASS: SYNCS.*, UBLKCP) -------------------------------------- */
__device__ __forceinline__ unsigned fep_smem_addr(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void fep_mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
/* try_wait suspends the thread for a hardware-chosen time slice; a copy that has not landed after 2 s of
 * them is not coming (bad source range): say so and trap instead of hanging the GPU.  bar = shared-window address. */
__device__ __forceinline__ bool fep_mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    return ok != 0;
}
static __device__ __noinline__ void fep_mbar_wait_slow(unsigned bar, unsigned parity, unsigned int* fault)
{
    unsigned long long t0 = 0;
    for (unsigned spins = 1; !fep_mbar_try_wait(bar, parity); spins++)
    {
        if ((spins & 255u) == 0u)
        {
            const unsigned long long now = fep_globaltimer();
            if (t0 == 0)
            {
                t0 = now;
            }
            else if (now - t0 > 2000000000ull)
            {
                fep_fault(fault, FEP_FAULT_STAGE_TIMEOUT, blockIdx.x, threadIdx.x, parity);
            }
        }
    }
}

/* one lane of the (converged) warp, chosen by the hardware: the compiler keeps what follows in uniform registers */
__device__ __forceinline__ bool fep_elect_one()
{
    unsigned pred;
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "elect.sync _|p, 0xffffffff;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(pred));
    return pred != 0;
}

/* ---- the trips of one warp ---------------------------------------------------------------------- */
/* A warp evaluates the runs  first, first + stride, ...  (first = its number in the grid, stride = warps of the grid),
 * each run = run_trips consecutive trips.  A cursor walks that sequence: `t` = trip it stands on, `left` = trips left
 * in the current run including t; t >= end: exhausted.  All of it is warp-uniform. */
struct FepWalk
{
    int run_trips; /* R */
    int jump;      /* trips between the starts of two consecutive runs of the warp */
    int end;       /* ka.trip_end */
};
struct FepCursor
{
    int t, left;
};
/* run_trips: ka.run_trips for a kernel that keeps per-segment sums, 1 for one that does not (any split of the trips
 * will do: single trips balance best) */
__device__ __forceinline__ FepWalk fep_walk(const KernelArgs& ka, int warps_of_grid, int run_trips)
{
    FepWalk w;
    w.run_trips = run_trips;
    w.jump      = warps_of_grid * run_trips;
    w.end       = ka.trip_end;
    return w;
}
__device__ __forceinline__ FepCursor fep_cursor(const KernelArgs& ka, const FepWalk& w, int warp_of_grid)
{
    FepCursor c;
    c.t    = ka.trip_begin + warp_of_grid * w.run_trips;
    c.left = w.run_trips;
    return c;
}
__device__ __forceinline__ bool fep_cursor_valid(const FepCursor& c, const FepWalk& w)
{
    return c.t < w.end;
}
__device__ __forceinline__ void fep_cursor_next(FepCursor& c, const FepWalk& w)
{
    c.t++;
    if (--c.left == 0)
    {
        c.t += w.jump - w.run_trips;
        c.left = w.run_trips;
    }
}

/* ---- a warp's ring of trip blocks in shared memory ---------------------------------------------- */
__host__ __device__ __forceinline__ size_t fep_ring_bytes(int warps)
{
    return (size_t)warps * FEP_RING_DEPTH * FEP_TRIP_WORDS * sizeof(unsigned int);
}

/* Ring slot q % FEP_RING_DEPTH holds the warp's q-th trip; the slot's mbarrier completes phase q / FEP_RING_DEPTH
 * when the block has landed.  The same warp fills and drains its ring, so "slot free" needs no second barrier:
 * the copy into a slot is issued after the warp has converged behind its last read of that slot. */
template<bool STAGED>
struct FepRing
{
    const unsigned int* trips;  /* global */
    const unsigned int* smem;   /* this warp's FEP_RING_DEPTH blocks */
    unsigned            smem_s; /* ... as a shared-window address */
    unsigned            bars_s; /* this warp's FEP_RING_DEPTH mbarriers, shared-window address */
    unsigned int*       fault;
};

/* one elected lane: expect + bulk copy of trip t into slot */
template<bool STAGED>
__device__ __forceinline__ void fep_ring_issue(const FepRing<STAGED>& r, int slot, int t)
{
    if (STAGED && fep_elect_one())
    {
        const unsigned bar = r.bars_s + 8u * (unsigned)slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(FEP_TRIP_WORDS * 4u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             r.smem_s + (unsigned)slot * (FEP_TRIP_WORDS * 4u)),
                     "l"(r.trips + (size_t)t * FEP_TRIP_WORDS), "r"(FEP_TRIP_WORDS * 4u), "r"(bar)
                     : "memory");
    }
}

/* `smem`: the CTA's dynamic shared memory (fep_ring_bytes), `bars`: FEP_RING_DEPTH mbarriers per warp.  Starts the
 * copies of the warp's first FEP_RING_DEPTH - 1 trips and leaves `issue` on the next trip to bring in.  Called by all
 * lanes of the warp. */
template<bool STAGED>
__device__ __forceinline__ FepRing<STAGED> fep_ring_open(const KernelArgs& ka, const FepWalk& w, FepCursor& issue,
                                                         unsigned char* smem, unsigned long long* bars, int warp)
{
    FepRing<STAGED> r;
    r.trips  = ka.trips;
    r.fault  = ka.fault;
    r.smem   = reinterpret_cast<const unsigned int*>(smem) + (size_t)warp * FEP_RING_DEPTH * FEP_TRIP_WORDS;
    r.smem_s = fep_smem_addr(r.smem);
    r.bars_s = fep_smem_addr(bars + warp * FEP_RING_DEPTH);
    if (STAGED)
    {
        if (fep_elect_one())
        {
#pragma unroll
            for (int s = 0; s < FEP_RING_DEPTH; s++)
            {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(r.bars_s + 8u * s) : "memory");
            }
            fep_mbar_init_fence();
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < FEP_RING_DEPTH - 1; q++)
        {
            if (fep_cursor_valid(issue, w))
            {
                fep_ring_issue<STAGED>(r, q, issue.t);
            }
            fep_cursor_next(issue, w);
        }
    }
    return r;
}

/* the block of the warp's q-th trip (trip number t): waits until it has landed; STAGED = false (the A/B variant
 * without shared memory, profiles/) reads the block where it lies in global memory */
template<bool STAGED>
__device__ __forceinline__ const unsigned int* fep_ring_block(const FepRing<STAGED>& r, int q, int t)
{
    if (STAGED)
    {
        const int      s      = q & (FEP_RING_DEPTH - 1);
        const unsigned parity = ((unsigned)q / FEP_RING_DEPTH) & 1u;
        const unsigned bar    = r.bars_s + 8u * (unsigned)s;
        if (!fep_mbar_try_wait(bar, parity))
        {
            fep_mbar_wait_slow(bar, parity, r.fault);
        }
        return r.smem + s * FEP_TRIP_WORDS;
    }
    return r.trips + (size_t)t * FEP_TRIP_WORDS;
}

template<bool STAGED>
__device__ __forceinline__ unsigned int fep_tw(const unsigned int* p)
{
    return STAGED ? *p : __ldg(p);
}

/* ---- per trip / per slot ------------------------------------------------------------------------ */
/* what a lane has in flight for the NEXT trip: the words that address its gathers, and the gathers */
struct FepFetch
{
    unsigned int head; /* owner | shift_eff << 24 | flipped << 30 */
    unsigned int cjx;
    float3       xo, xj; /* owner and partner coordinates */
    float3       sh;     /* shift vector of the trip */
    float4       ta, tb; /* type-table rows of states A and B: {c6, c12, sigma6, c6grid} */
};

/* a row of the type table; without LJ-PME the fourth word (c6grid) is not loaded: an unused destination register
 * of a load in flight is handed out again by the compiler and stalls its next writer until the load has landed */
template<bool GRID>
__device__ __forceinline__ float4 fep_load_type_row(const float4* __restrict__ typetab, unsigned int row)
{
    if (GRID)
    {
        return __ldg(typetab + row);
    }
    const float2 a = __ldg(reinterpret_cast<const float2*>(typetab + row));
    const float  z = __ldg(reinterpret_cast<const float*>(typetab + row) + 2);
    return make_float4(a.x, a.y, z, 0.0f);
}

/* tb = first word of the trip's block; GRID: the kernel uses c6grid (LJ-PME) */
template<bool STAGED, bool GRID = true>
__device__ __forceinline__ FepFetch fep_fetch(const KernelArgs& ka, const unsigned int* tb, int lane)
{
    FepFetch f;
    f.head = fep_tw<STAGED>(tb + FEP_TH_OWNER);
    f.cjx  = fep_tw<STAGED>(tb + FEP_TW_CJX + lane);
    /* nbfp row = type of the reference's i atom (:499-500), column = type of its j atom (:560-563): the list
     * builder has resolved orientation and owner type into the two table indices of the pair */
    const unsigned int tt = fep_tw<STAGED>(tb + FEP_TW_TJ + lane);
    f.xo                  = fep_load_pos(ka.pos3, (int)(f.head & (FEP_MAX_TOUCHED - 1)));
    f.xj                  = fep_load_pos(ka.pos3, (int)(f.cjx & (FEP_MAX_TOUCHED - 1)));
    /* three scalars, not one float4: an unused fourth register is handed out again while the load is in flight */
    const float* sv = reinterpret_cast<const float*>(ka.dyn->shiftvec + ((f.head >> 24) & 63u));
    f.sh            = make_float3(__ldg(sv), __ldg(sv + 1), __ldg(sv + 2));
    f.ta            = fep_load_type_row<GRID>(ka.typetab, tt & 0xffffu);
    f.tb            = fep_load_type_row<GRID>(ka.typetab, tt >> 16);
    return f;
}

/* everything lambda-independent about the pair in this lane's slot */
struct FepSlot
{
    bool   active, excluded, self, within, contrib;
    float  dx, dy, dz, r2; /* r2 not yet clamped */
    float  qq[2];          /* NOT masked: lanes without a contributing pair must be masked by the caller */
    float4 ta, tb;         /* type-table rows of states A and B: {c6, c12, sigma6, c6grid} */
};

template<bool STAGED>
__device__ __forceinline__ FepSlot fep_slot(const KernelArgs& ka, const unsigned int* tb, const FepFetch& f, int lane)
{
    FepSlot p;
    p.ta = f.ta;
    p.tb = f.tb;
    const unsigned int owner = f.head & (FEP_MAX_TOUCHED - 1);
    const unsigned int cj    = f.cjx & (FEP_MAX_TOUCHED - 1);
    p.active                 = (f.cjx & FEP_SLOT_PADDING) == 0;
    p.excluded               = (int)f.cjx < 0;
    p.self                   = owner == cj;
    /* the reference shifts the i atom first (:478-480); here the owner plays that part */
    p.dx = (f.sh.x + f.xo.x) - f.xj.x;
    p.dy = (f.sh.y + f.xo.y) - f.xj.y;
    p.dz = (f.sh.z + f.xo.z) - f.xj.z;
    p.r2 = fmaf(p.dz, p.dz, fmaf(p.dy, p.dy, p.dx * p.dx));
    p.within  = p.r2 < ka.rcut_max2;
    p.contrib = p.active && (p.within || p.excluded); /* :667 */
    const float qoA = __uint_as_float(fep_tw<STAGED>(tb + FEP_TH_QA)), qoB = __uint_as_float(fep_tw<STAGED>(tb + FEP_TH_QB));
    const float qjA = __uint_as_float(fep_tw<STAGED>(tb + FEP_TW_QA + lane));
    const float qjB = __uint_as_float(fep_tw<STAGED>(tb + FEP_TW_QB + lane));
    p.qq[0]         = (ka.epsfac * qoA) * qjA;
    p.qq[1]         = (ka.epsfac * qoB) * qjB;
    return p;
}

/* The generic kernels' description of the pair (fep_pair_math.cuh) from the slot: everything that does not
 * depend on lambda.  Returns false when the slot contributes nothing (padding, or included and beyond the
 * cut-off sphere, reference :667). */
template<int SC>
__device__ __forceinline__ bool fep_fill_pair(const KernelArgs& ka, const FepSlot& p, FepPair& pr)
{
    if (!p.contrib)
    {
        return false;
    }
    const float4 a = p.ta, b = p.tb;
    pr.qq[0] = p.qq[0], pr.qq[1] = p.qq[1];
    pr.c6[0] = a.x, pr.c12[0] = a.y, pr.sig6[0] = a.z, pr.c6g[0] = a.w;
    pr.c6[1] = b.x, pr.c12[1] = b.y, pr.sig6[1] = b.z, pr.c6g[1] = b.w;
    /* soft-core only if one end state has no repulsion (:597-628) */
    const bool hard = (a.y > 0.0f && b.y > 0.0f);
    if (SC == FEP_SC_BEUTLER)
    {
        pr.a_c = hard ? 0.0f : ka.alpha_c;
        pr.a_v = hard ? 0.0f : ka.alpha_v;
    }
    else if (SC == FEP_SC_GAPSYS)
    {
        pr.a_c      = hard ? 0.0f : ka.gscale_c;
        pr.a_v      = hard ? 0.0f : ka.gscale_v;
        pr.gbase[0] = fep_sixth_root((26.0f / 7.0f) * a.z);
        pr.gbase[1] = fep_sixth_root((26.0f / 7.0f) * b.z);
    }
    else
    {
        pr.a_c = pr.a_v = 0.0f;
    }
    const float r2 = fmaxf(p.r2, FEP_MIN_RSQ);
    pr.r2   = r2;
    pr.rinv = fep_rsqrt(r2);
    pr.r    = r2 * pr.rinv;
    if (SC == FEP_SC_BEUTLER)
    {
        pr.rpm2 = r2 * r2;
        pr.r6   = pr.rpm2 * r2;
    }
    else
    {
        pr.rpm2 = pr.rinv * pr.rinv;
    }
    pr.nonzero[0]      = (pr.qq[0] != 0.0f || a.x != 0.0f || a.y != 0.0f);
    pr.nonzero[1]      = (pr.qq[1] != 0.0f || b.x != 0.0f || b.y != 0.0f);
    pr.included_within = p.within && !p.excluded;
    return true;
}

/* sum of v over the warp, valid in every lane */
__device__ __forceinline__ float fep_warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        v += __shfl_xor_sync(FEP_FULL_MASK, v, o);
    }
    return v;
}

/* The owner's side of a segment (fep_types.h): per-lane sums over the segment's trips, reduced and stored by
 * the segment's last trip.  f = MINUS the force on the owner (= what the partners received), vc / vv = the
 * segment's energies (kept only with several energy-group pairs). */
struct FepSegment
{
    float fx, fy, fz, vc, vv;
};
__device__ __forceinline__ void fep_segment_clear(FepSegment& s)
{
    s.fx = s.fy = s.fz = s.vc = s.vv = 0.0f;
}
/* tb: block of the segment's last trip; head: its header word 0 */
template<bool STAGED>
__device__ __forceinline__ void fep_segment_flush(const KernelArgs& ka, const unsigned int* tb, unsigned int head,
                                                  FepSegment& s, bool want_shift, bool per_segment_energy, int lane)
{
    const float fx = fep_warp_sum(s.fx), fy = fep_warp_sum(s.fy), fz = fep_warp_sum(s.fz);
    float       vc = 0.0f, vv = 0.0f;
    if (per_segment_energy)
    {
        vc = fep_warp_sum(s.vc);
        vv = fep_warp_sum(s.vv);
    }
    if (lane == 0)
    {
        ka.fsorted[fep_tw<STAGED>(tb + FEP_TH_SLOT_F)] = make_float4(-fx, -fy, -fz, 0.0f);
        if (want_shift)
        {
            /* nb_free_energy.cpp:1153-1164 adds the i atom's force to the entry's shift vector; for a flipped
             * trip the owner was the j atom, whose force is minus that */
            const float sg = (head & FEP_TRIP_FLIPPED) ? 1.0f : -1.0f;
            ka.fshift_sorted[fep_tw<STAGED>(tb + FEP_TH_SLOT_SHIFT)] = make_float4(sg * fx, sg * fy, sg * fz, 0.0f);
        }
        if (per_segment_energy)
        {
            ka.ev2[fep_tw<STAGED>(tb + FEP_TH_SLOT_EV)] = make_float2(vc, vv);
        }
    }
    fep_segment_clear(s);
}

#endif
