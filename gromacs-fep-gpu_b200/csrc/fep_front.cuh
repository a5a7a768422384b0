/*
 * fep_front.cuh -- how the pair kernels get their input: the front end shared by fep_beutler.cu and
 * fep_kernels.cu on the trip layout of fep_types.h.
 *
 *   - A CTA owns a contiguous tile of trips.  The per-slot records of the tile (partner index, scatter
 *     destination, partner charges, partner types) and the trip records are contiguous in global
 *     memory (one block of FEP_TRIP_WORDS words per trip), so the whole tile is brought into shared
 *     memory by ONE bulk copy (cp.async.bulk, completion on an mbarrier) issued by one thread before anything else happens:
 *     the list stream costs no registers, no address arithmetic and no load instructions in the loop.
 *   - Per trip the warp-uniform data (owner coordinates + shift vector; owner charges and owner rows of
 *     the type table, pre-gathered into the trip's header) replace what the reference sets up per
 *     i-entry (nb_free_energy.cpp:466-503).
 *   - The only dependent per-pair load is the partner's coordinates; it is issued one trip ahead.
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
__device__ __forceinline__ void fep_mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fep_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fep_mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fep_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fep_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         fep_smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(fep_smem_addr(bar))
                 : "memory");
}
/* try_wait suspends the thread for a hardware-chosen time slice; a copy that has not landed after 2 s of
 * them is not coming (bad source range): say so and trap instead of hanging the GPU */
__device__ __forceinline__ bool fep_mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(fep_smem_addr(bar)), "r"(parity)
            : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fep_mbar_wait(unsigned long long* bar, unsigned parity, unsigned int* fault)
{
    if (fep_mbar_try_wait(bar, parity))
    {
        return;
    }
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

/* ---- the tile of one CTA in shared memory ------------------------------------------------------- */
/* trip blocks (fep_types.h) are contiguous: a tile of nt trips is one copy of nt * FEP_TRIP_WORDS words */
__host__ __device__ __forceinline__ size_t fep_tile_bytes(int tile_trips)
{
    return (size_t)tile_trips * FEP_TRIP_WORDS * sizeof(unsigned int);
}

/* The tile arrives in up to FEP_STAGE_CHUNKS pieces, each a whole number of rounds (a round = one trip per
 * warp of the CTA), each with its own mbarrier: the first round can start as soon as the first piece is in
 * shared memory while the rest of the tile is still on its way. */
#define FEP_STAGE_CHUNKS 8

struct FepStage
{
    const unsigned int* tile;        /* first word of the tile (shared memory, or global when not staged) */
    int                 chunk_trips; /* trips per piece */
};

/* Stages trips [t0, t0 + nt) (nt >= 1).  STAGED = false is the A/B variant without shared memory: the
 * "tile" then points at global memory (profiles/: staged vs direct).  `bars` (FEP_STAGE_CHUNKS mbarriers)
 * and `smem` are the CTA's; every thread of the CTA must call this, then __syncthreads() (which publishes
 * the barriers' initialisation); a warp calls fep_stage_wait() before the first trip of every piece. */
template<bool STAGED>
__device__ __forceinline__ FepStage fep_stage_tile(const KernelArgs& ka, int t0, int nt, int warps, unsigned char* smem,
                                                   unsigned long long* bars)
{
    FepStage            st;
    const unsigned int* src = ka.trips + (size_t)t0 * FEP_TRIP_WORDS;
    const int           rounds = (nt + warps - 1) / warps;
    st.chunk_trips             = warps * ((rounds + FEP_STAGE_CHUNKS - 1) / FEP_STAGE_CHUNKS);
    if (!STAGED)
    {
        st.tile = src;
        return st;
    }
    if (threadIdx.x == 0)
    {
        for (int c = 0; c < FEP_STAGE_CHUNKS; c++)
        {
            fep_mbar_init(bars + c, 1);
        }
        for (int c = 0, b = 0; b < nt; c++, b += st.chunk_trips)
        {
            const unsigned bytes = (unsigned)min(st.chunk_trips, nt - b) * FEP_TRIP_WORDS * 4u;
            fep_mbar_expect_tx(bars + c, bytes);
            fep_bulk_g2s(smem + (size_t)b * FEP_TRIP_WORDS * 4u, src + (size_t)b * FEP_TRIP_WORDS, bytes, bars + c);
        }
    }
    st.tile = reinterpret_cast<const unsigned int*>(smem);
    return st;
}

/* before local trip lt is read: waits for its piece when lt is the warp's first trip in it */
template<bool STAGED>
__device__ __forceinline__ void fep_stage_wait(const FepStage& st, unsigned long long* bars, int lt, int warps,
                                               unsigned int* fault)
{
    if (STAGED)
    {
        const int c = lt / st.chunk_trips;
        if (lt - c * st.chunk_trips < warps)
        {
            fep_mbar_wait(bars + c, 0, fault);
        }
    }
}

template<bool STAGED>
__device__ __forceinline__ unsigned int fep_tw(const unsigned int* p)
{
    return STAGED ? *p : __ldg(p);
}

/* ---- per trip / per slot ------------------------------------------------------------------------ */
/* what a lane has in flight for the NEXT trip: its record and the two dependent gathers */
struct FepFetch
{
    unsigned int head; /* owner | shift_eff << 24 | flipped << 30 */
    unsigned int cjx;
    float3       xo, xj; /* owner and partner coordinates */
    float4       ta, tb; /* type-table rows of states A and B: {c6, c12, sigma6, c6grid} */
};

/* tb = first word of the trip's block */
template<bool STAGED>
__device__ __forceinline__ FepFetch fep_fetch(const KernelArgs& ka, const unsigned int* tb, int lane)
{
    FepFetch f;
    f.head = fep_tw<STAGED>(tb + FEP_TH_OWNER);
    f.cjx  = fep_tw<STAGED>(tb + FEP_TW_CJX + lane);
    f.xo   = fep_load_pos(ka.pos3, (int)(f.head & (FEP_MAX_TOUCHED - 1)));
    f.xj   = fep_load_pos(ka.pos3, (int)(f.cjx & (FEP_MAX_TOUCHED - 1)));
    /* nbfp row = type of the reference's i atom (:499-500), column = type of its j atom (:560-563): for a
     * flipped trip the partner is the i atom */
    const unsigned int tt  = fep_tw<STAGED>(tb + FEP_TW_TJ + lane);
    const int          mul = (f.head & FEP_TRIP_FLIPPED) ? ka.ntype : 1;
    const int          iA  = mul * (int)(tt & 0xffffu) + (int)fep_tw<STAGED>(tb + FEP_TH_TADD_A);
    const int          iB  = mul * (int)(tt >> 16) + (int)fep_tw<STAGED>(tb + FEP_TH_TADD_B);
    f.ta                   = __ldg(ka.typetab + iA);
    f.tb                   = __ldg(ka.typetab + iB);
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
__device__ __forceinline__ FepSlot fep_slot(const KernelArgs& ka, const unsigned int* tb, const FepFetch& f, int lane,
                                            const float4* s_shift)
{
    FepSlot            p;
    const unsigned int owner = f.head & (FEP_MAX_TOUCHED - 1);
    const unsigned int cj    = f.cjx & (FEP_MAX_TOUCHED - 1);
    p.active                 = (f.cjx & FEP_SLOT_PADDING) == 0;
    p.excluded               = (int)f.cjx < 0;
    p.self                   = owner == cj;
    const float4 sh          = s_shift[min((f.head >> 24) & 63u, (unsigned)(FEP_NUM_SHIFT - 1))];
    /* the reference shifts the i atom first (:478-480); here the owner plays that part */
    p.dx = (sh.x + f.xo.x) - f.xj.x;
    p.dy = (sh.y + f.xo.y) - f.xj.y;
    p.dz = (sh.z + f.xo.z) - f.xj.z;
    p.r2 = fmaf(p.dz, p.dz, fmaf(p.dy, p.dy, p.dx * p.dx));
    p.within  = p.r2 < ka.rcut_max2;
    p.contrib = p.active && (p.within || p.excluded); /* :667 */
    const float qoA = __uint_as_float(fep_tw<STAGED>(tb + FEP_TH_QA)), qoB = __uint_as_float(fep_tw<STAGED>(tb + FEP_TH_QB));
    const float qjA = __uint_as_float(fep_tw<STAGED>(tb + FEP_TW_QA + lane));
    const float qjB = __uint_as_float(fep_tw<STAGED>(tb + FEP_TW_QB + lane));
    p.qq[0]         = (ka.epsfac * qoA) * qjA;
    p.qq[1]         = (ka.epsfac * qoB) * qjB;
    p.ta            = f.ta;
    p.tb            = f.tb;
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

#endif
