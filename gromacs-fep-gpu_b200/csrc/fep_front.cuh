/*
 * fep_front.cuh -- how the pair kernels get their input: the front end shared by fep_beutler.cu and
 * fep_kernels.cu on the trip layout of fep_types.h.
 *
 *   - A CTA owns a contiguous tile of trips.  The per-slot records of the tile (partner index, scatter
 *     destination, partner charges, partner types) and the trip records are contiguous in global
 *     memory, so the whole tile is brought into shared memory by at most five bulk copies
 *     (cp.async.bulk, completion on an mbarrier) issued by one thread before anything else happens:
 *     the list stream costs no registers, no address arithmetic and no load instructions in the loop.
 *   - Per trip the warp-uniform data (owner coordinates + shift vector, owner charges, owner rows of
 *     the type table) replace what the reference sets up per i-entry (nb_free_energy.cpp:466-503).
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
__device__ __forceinline__ void fep_mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "FEP_WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra FEP_DONE_%=;\n"
            "bra FEP_WAIT_%=;\n"
            "FEP_DONE_%=:\n"
            "}\n" ::"r"(fep_smem_addr(bar)),
            "r"(parity)
            : "memory");
}

/* ---- the tile of one CTA in shared memory ------------------------------------------------------- */
/* dynamic shared memory: [trip4: tt int4][cjx: 32 tt int][qj: 32 tt float2][tj: 32 tt int]([dst: 32 tt int]) */
struct FepTile
{
    const int4*   trip4;
    const int*    cjx;
    const float2* qj;
    const int*    tj;
    const int*    dst;
};

__host__ __device__ __forceinline__ size_t fep_tile_bytes(int tile_trips, bool with_dst)
{
    return (size_t)tile_trips * (sizeof(int4) + 32 * (sizeof(int) + sizeof(float2) + sizeof(int) + (with_dst ? sizeof(int) : 0)));
}

/* Stages trips [t0, t0 + nt) (nt >= 1).  STAGED = false is the A/B variant without shared memory: the
 * "tile" then points at global memory (profiles/: staged vs direct).  `bar` and `smem` are the CTA's
 * mbarrier and dynamic shared memory; every thread of the CTA must call this. */
template<bool STAGED, bool WITH_DST>
__device__ __forceinline__ FepTile fep_stage_tile(const KernelArgs& ka, int t0, int nt, int tile_trips, unsigned char* smem,
                                                  unsigned long long* bar)
{
    FepTile t;
    if (!STAGED)
    {
        t.trip4 = ka.trip4 + t0;
        t.cjx   = ka.cjx + 32 * (size_t)t0;
        t.qj    = ka.qj + 32 * (size_t)t0;
        t.tj    = ka.tj + 32 * (size_t)t0;
        t.dst   = ka.dst + 32 * (size_t)t0;
        return t;
    }
    int4*   s_trip4 = reinterpret_cast<int4*>(smem);
    int*    s_cjx   = reinterpret_cast<int*>(s_trip4 + tile_trips);
    float2* s_qj    = reinterpret_cast<float2*>(s_cjx + 32 * tile_trips);
    int*    s_tj    = reinterpret_cast<int*>(s_qj + 32 * tile_trips);
    int*    s_dst   = s_tj + 32 * tile_trips;
    if (threadIdx.x == 0)
    {
        fep_mbar_init(bar, 1);
        const unsigned n4 = (unsigned)nt * 32u * 4u, n8 = 2u * n4, n16 = (unsigned)nt * 16u;
        fep_mbar_expect_tx(bar, n16 + 2u * n4 + n8 + (WITH_DST ? n4 : 0u));
        fep_bulk_g2s(s_trip4, ka.trip4 + t0, n16, bar);
        fep_bulk_g2s(s_cjx, ka.cjx + 32 * (size_t)t0, n4, bar);
        fep_bulk_g2s(s_qj, ka.qj + 32 * (size_t)t0, n8, bar);
        fep_bulk_g2s(s_tj, ka.tj + 32 * (size_t)t0, n4, bar);
        if (WITH_DST)
        {
            fep_bulk_g2s(s_dst, ka.dst + 32 * (size_t)t0, n4, bar);
        }
    }
    t.trip4 = s_trip4;
    t.cjx   = s_cjx;
    t.qj    = s_qj;
    t.tj    = s_tj;
    t.dst   = s_dst;
    return t;
}

/* call after fep_stage_tile() and a __syncthreads() that makes the barrier's initialisation visible */
template<bool STAGED>
__device__ __forceinline__ void fep_tile_wait(unsigned long long* bar)
{
    if (STAGED)
    {
        fep_mbar_wait(bar, 0);
    }
}

/* ---- per trip / per slot ------------------------------------------------------------------------ */
/* what a lane has in flight for the NEXT trip: its record and the three dependent gathers */
struct FepFetch
{
    int    td_x;   /* trip4.x */
    int    cjx;
    float3 xo, xj; /* owner and partner coordinates */
    float4 po;     /* owner parameters */
};

template<bool STAGED>
__device__ __forceinline__ FepFetch fep_fetch(const KernelArgs& ka, const FepTile& tile, int lt, int lane)
{
    FepFetch f;
    f.td_x = STAGED ? tile.trip4[lt].x : __ldg(&tile.trip4[lt].x);
    f.cjx  = STAGED ? tile.cjx[32 * lt + lane] : __ldg(tile.cjx + 32 * lt + lane);
    const int owner = f.td_x & (FEP_MAX_TOUCHED - 1);
    f.xo   = fep_load_pos(ka.pos3, owner);
    f.po   = __ldg(ka.par4 + owner);
    f.xj   = fep_load_pos(ka.pos3, f.cjx & (FEP_MAX_TOUCHED - 1));
    return f;
}

/* everything lambda-independent about the pair in this lane's slot */
struct FepSlot
{
    bool   active, excluded, self, within, contrib;
    float  dx, dy, dz, r2; /* r2 not yet clamped */
    float  qq[2];
    float4 ta, tb;         /* type-table rows of states A and B: {c6, c12, sigma6, c6grid} */
};

template<bool STAGED>
__device__ __forceinline__ FepSlot fep_slot(const KernelArgs& ka, const FepTile& tile, const FepFetch& f, int lt, int lane,
                                            const float4* s_shift)
{
    FepSlot   p;
    const int owner = f.td_x & (FEP_MAX_TOUCHED - 1);
    const int cj    = f.cjx & (FEP_MAX_TOUCHED - 1);
    const bool flip = (f.td_x & FEP_TRIP_FLIPPED) != 0;
    p.active        = (f.cjx & FEP_SLOT_PADDING) == 0;
    p.excluded      = f.cjx < 0;
    p.self          = owner == cj;
    const float4 sh = s_shift[min((f.td_x >> 24) & 63, FEP_NUM_SHIFT - 1)];
    /* the reference shifts the i atom first (:478-480); here the owner plays that part */
    p.dx = (sh.x + f.xo.x) - f.xj.x;
    p.dy = (sh.y + f.xo.y) - f.xj.y;
    p.dz = (sh.z + f.xo.z) - f.xj.z;
    p.r2 = fmaf(p.dz, p.dz, fmaf(p.dy, p.dy, p.dx * p.dx));
    p.within  = p.r2 < ka.rcut_max2;
    p.contrib = p.active && (p.within || p.excluded); /* :667 */
    const float2 q  = STAGED ? tile.qj[32 * lt + lane] : __ldg(tile.qj + 32 * lt + lane);
    const int    tt = STAGED ? tile.tj[32 * lt + lane] : __ldg(tile.tj + 32 * lt + lane);
    const float  m  = p.contrib ? 1.0f : 0.0f;
    p.qq[0]         = (ka.epsfac * f.po.x) * q.x * m;
    p.qq[1]         = (ka.epsfac * f.po.y) * q.y * m;
    /* nbfp row = type of the reference's i atom (:499-500), column = type of its j atom (:560-563) */
    const int toA = __float_as_int(f.po.z), toB = __float_as_int(f.po.w);
    const int tjA = tt & 0xffff, tjB = (tt >> 16) & 0xffff;
    const int iA  = flip ? ka.ntype * tjA + toA : ka.ntype * toA + tjA;
    const int iB  = flip ? ka.ntype * tjB + toB : ka.ntype * toB + tjB;
    p.ta          = __ldg(ka.typetab + iA);
    p.tb          = __ldg(ka.typetab + iB);
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

/* the whole list as one "tile" in global memory (kernels that do not stage) */
__device__ __forceinline__ FepTile fep_global_tile(const KernelArgs& ka)
{
    FepTile t;
    t.trip4 = ka.trip4;
    t.cjx   = ka.cjx;
    t.qj    = ka.qj;
    t.tj    = ka.tj;
    t.dst   = ka.dst;
    return t;
}

#endif
