/*
 * fep_gapsys.cu -- the energy-only foreign-lambda passes of the Gapsys soft-core (BASELINE.json configuration 3) with
 * everything lambda-independent hoisted out of the loop over lambda points.
 *
 * Reference: nbnxm/freeenergydispatch.cpp:236-306 calls the energy-only flavour of nb_free_energy_kernel<Gapsys>
 * (gmxlib/nonbonded/nb_free_energy.cpp:747-1020 with nb_softcore.h:73-279) once per lambda point.
 * The Gapsys soft-core replaces, for r below a linearisation radius rQ, the plain Coulomb / Lennard-Jones energy of a
 * state by a quadratic in r whose coefficients depend on rQ, and
 *     rQ_coul(s, p) = (1 - lfacC[s][p])^(1/6) (1 + |qq_s| / facel) scale_c,   rQ_vdw(s, p) = (1 - lfacV[s][p])^(1/6) (26/7 sigma6_s)^(1/6) scale_v
 * is the ONLY way lambda enters a state's energy.  So per pair and state the plain energies are evaluated once, and a
 * lambda point costs something only for pairs inside the LARGEST radius any lambda can give (the factor is <= 1):
 *     E(p)         = sum_s lfacC[s][p] (C_s + dC_s[p]) + lfacV[s][p] (G_s + dG_s[p])
 *     dVdl_coul(p) = (C_B + dC_B[p]) - (C_A + dC_A[p]) + X_coul[p],      dVdl_vdw(p) likewise
 * with C_s, G_s = sums over pairs of the plain state energies plus the real-space corrections (:1023-1136), dC, dG =
 * sums of (soft-core form - plain form) over the pairs inside rQ(s, p), X = the explicit d/dlambda terms of the
 * linearisation (nb_softcore.h:167-176, 263-271).  State-A sums and per-pair B-minus-A differences are accumulated,
 * like in fep_beutler_kernel.  A warp in which no lane is inside the largest radius skips the point loop altogether
 * (C3: cut-off 1.0 nm against radii of ~0.3 nm).  The per-pair formulas are those of fep_included_terms<GAPSYS>
 * (fep_pair_math.cuh), which stays the reference for forces and for potential-switch runs.
 *
 * Input side: the trip layout, per-warp rings and walk of fep_front.cuh; single trips are dealt to the warps.
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fep_front.cuh"

#define FEP_GS_MAXC 8 /* lambda points per launch */
#define FEP_GS_CTA 128

struct GapsysStep
{
    int   n_tiles;                /* CTAs per chunk = stride of the rows of for_part (gridDim.x) */
    float g6c_max[2], g6v_max[2]; /* largest (1 - lfac)^(1/6) of all points per state (1 unless a lambda < 0) */
};

extern __shared__ __align__(128) unsigned char fep_gs_smem[];

/* per point: dG_A, D(dG), dC_A, D(dC), X_coul, X_vdw; then C_A, DC, G_A, DG */
template<int C>
struct GsLayout
{
    static constexpr int NACC = 6 * C + 4;
    static constexpr int N8   = (NACC + 7) / 8;
    static constexpr int iVA = 0, iDV = C, iCA = 2 * C, iDCp = 3 * C, iXC = 4 * C, iXV = 5 * C;
    static constexpr int iC = 6 * C, iDC = iC + 1, iG = iC + 2, iDG = iC + 3;
};

template<bool EWALD, int C>
__global__ void __launch_bounds__(FEP_GS_CTA, (C > 4 ? 3 : 4))
        fep_gapsys_foreign_kernel(const __grid_constant__ KernelArgs ka, const __grid_constant__ GapsysStep gs)
{
    using L           = GsLayout<C>;
    constexpr int NW  = FEP_GS_CTA / 32;
    __shared__ float  s_red[NW][L::N8 * 8];
    __shared__ double s_sum[L::N8 * 8];
    __shared__ __align__(8) unsigned long long s_bars[NW * FEP_RING_DEPTH];

    __shared__ LambdaPoint s_pts[C]; /* the lambda points of this chunk (blockIdx.y) */

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(FEP_FULL_MASK, tid >> 5, 0);
    const int p0   = blockIdx.y * C;
    const int np   = min(C, ka.n_points - p0);
    fep_pdl_launch_dependents();
    for (int i = tid; i < C * (int)(sizeof(LambdaPoint) / 4); i += FEP_GS_CTA)
    {
        /* padding points repeat the last valid one; their results are not written */
        const int p = min(i / (int)(sizeof(LambdaPoint) / 4), np - 1), w = i % (int)(sizeof(LambdaPoint) / 4);
        reinterpret_cast<int*>(s_pts)[i] = __ldg(reinterpret_cast<const int*>(ka.pts + p0 + p) + w);
    }
    __syncthreads();

    const FepWalk       walk  = fep_walk(ka, gridDim.x * NW, 1);
    FepCursor           issue = fep_cursor(ka, walk, warp * gridDim.x + blockIdx.x);
    FepCursor           ahead = issue;
    const FepRing<true> ring  = fep_ring_open<true>(ka, walk, issue, fep_gs_smem, s_bars, warp);

    float acc[L::N8 * 8];
#pragma unroll
    for (int i = 0; i < L::N8 * 8; i++)
    {
        acc[i] = 0.0f;
    }

    const unsigned int* tb = nullptr;
    FepFetch            nx;
    bool                more = fep_cursor_valid(ahead, walk);
    if (more)
    {
        tb = fep_ring_block<true>(ring, 0, ahead.t);
        nx = fep_fetch<true>(ka, tb, lane);
        fep_cursor_next(ahead, walk);
    }
    for (int q = 0; more; q++)
    {
        const unsigned int* tb_cur = tb;
        const FepFetch      cur    = nx;
        __syncwarp();
        if (fep_cursor_valid(issue, walk))
        {
            fep_ring_issue<true>(ring, (q + FEP_RING_DEPTH - 1) & (FEP_RING_DEPTH - 1), issue.t);
        }
        fep_cursor_next(issue, walk);
        more = fep_cursor_valid(ahead, walk);
        if (more)
        {
            tb = fep_ring_block<true>(ring, q + 1, ahead.t);
            nx = fep_fetch<true>(ka, tb, lane);
            fep_cursor_next(ahead, walk);
        }
        const FepSlot sl = fep_slot<true>(ka, tb_cur, cur, lane);
        FepPair       pr;
        const bool    has = fep_fill_pair<FEP_SC_GAPSYS>(ka, sl, pr);

        /* what does not depend on lambda: corrections + the plain energies of both states */
        bool  cand_c[2] = { false, false }, cand_v[2] = { false, false };
        float base_c[2] = { 0.0f, 0.0f }, base_v[2] = { 0.0f, 0.0f }, vvP[2] = { 0.0f, 0.0f }, vcP_keep[2] = { 0.0f, 0.0f };
        if (has)
        {
            float xc, fc, xv, fv;
            fep_corrections<EWALD, false>(ka, pr, sl.excluded, sl.self, xc, fc, xv, fv);
            const float cA = pr.qq[0] * xc, cB = pr.qq[1] * xc, gA = pr.c6g[0] * xv, gB = pr.c6g[1] * xv;
            float       vcP[2] = { 0.0f, 0.0f }, vsh[2] = { 0.0f, 0.0f };
            if (pr.included_within)
            {
#pragma unroll
                for (int s = 0; s < 2; s++)
                {
                    /* :804-874 and :880-971 without soft-core */
                    const bool elec = pr.r < ka.rcoulomb && pr.qq[s] != 0.0f;
                    if (elec)
                    {
                        vcP[s] = EWALD ? pr.qq[s] * (pr.rinv - ka.sh_ewald)
                                       : pr.qq[s] * (pr.rinv + ka.krf * pr.r * pr.r - ka.crf);
                        base_c[s] = (1.0f + fabsf(pr.qq[s] / ka.gapsys_facel)) * pr.a_c;
                        cand_c[s] = pr.a_c > 0.0f && ka.gapsys_facel != 0.0f
                                    && pr.r < fminf(base_c[s] * gs.g6c_max[s], ka.gapsys_rcoul);
                    }
                    const bool vdw = pr.r < ka.rvdw && (pr.c6[s] != 0.0f || pr.c12[s] != 0.0f);
                    if (vdw)
                    {
                        const float ri2   = pr.rinv * pr.rinv;
                        const float rinv6 = fminf(ri2 * ri2 * ri2, FEP_MAX_RINV6);
                        const float v6 = pr.c6[s] * rinv6, v12 = pr.c12[s] * rinv6 * rinv6;
                        vvP[s] = v12 * (1.0f / 12.0f) - v6 * (1.0f / 6.0f);
                        /* the shift constants are the same with and without soft-core (:916-932, nb_softcore.h:258) */
                        vsh[s] = (pr.c12[s] * ka.rep_cpot) * (1.0f / 12.0f) - (pr.c6[s] * ka.disp_cpot) * (1.0f / 6.0f);
                        if (ka.vdw_ewald)
                        {
                            vsh[s] += pr.c6g[s] * ka.sh_lj_ewald * (1.0f / 6.0f);
                        }
                        base_v[s] = pr.gbase[s] * pr.a_v;
                        cand_v[s] = pr.a_v > 0.0f && pr.r < base_v[s] * gs.g6v_max[s];
                    }
                }
            }
            /* A state that may be soft-cored at some lambda keeps its WHOLE energy in the per-point sums (plain or
             * linearised, decided per point): plain energy in the lambda-independent sum and (linearised - plain) in
             * the per-point sum would cancel ~1e9 kJ/mol of two overlapping atoms in fp32 */
            const float cpA = cand_c[0] ? 0.0f : vcP[0], cpB = cand_c[1] ? 0.0f : vcP[1];
            const float gpA = cand_v[0] ? 0.0f : vvP[0], gpB = cand_v[1] ? 0.0f : vvP[1];
            acc[L::iC] += cA + cpA;
            acc[L::iDC] += (cB - cA) + (cpB - cpA);
            acc[L::iG] += gA + (gpA + vsh[0]);
            acc[L::iDG] += (gB - gA) + ((gpB + vsh[1]) - (gpA + vsh[0]));
            vcP_keep[0] = vcP[0], vcP_keep[1] = vcP[1];
        }
        /* the point loop only for warps with a lane inside the largest linearisation radius of some state */
        if (__any_sync(FEP_FULL_MASK, cand_c[0] || cand_c[1] || cand_v[0] || cand_v[1]))
        {
            const float r = pr.r, r2 = pr.r2;
#pragma unroll
            for (int p = 0; p < C; p++)
            {
                const LambdaPoint& lp = s_pts[p];
                float dc[2] = { 0.0f, 0.0f }, dv[2] = { 0.0f, 0.0f }, xc2[2] = { 0.0f, 0.0f }, xv2[2] = { 0.0f, 0.0f };
#pragma unroll
                for (int s = 0; s < 2; s++)
                {
                    /* Coulomb, nb_softcore.h:73-195 */
                    if (cand_c[s])
                    {
                        dc[s] = vcP_keep[s];
                        if (lp.lfac_c[s] < 1.0f)
                        {
                            float      rq         = lp.g6_c[s] * base_c[s];
                            const bool within_cut = rq <= ka.gapsys_rcoul;
                            rq                    = fminf(rq, ka.gapsys_rcoul);
                            if (r < rq)
                            {
                                const float rinvq = fep_rcp(rq);
                                const float cst   = pr.qq[s] * rinvq;
                                const float lin   = cst * r * rinvq;
                                const float quad  = lin * r * rinvq;
                                /* the shift / reaction-field part is the same in both forms */
                                dc[s]  = (quad - 3.0f * (lin - cst)) + (vcP_keep[s] - pr.qq[s] * pr.rinv);
                                xc2[s] = within_cut ? 0.5f * lp.gdl_c[s] * (quad - 2.0f * lin + cst) : 0.0f;
                            }
                        }
                    }
                    /* Lennard-Jones, nb_softcore.h:199-279 */
                    if (cand_v[s])
                    {
                        dv[s] = vvP[s];
                        if (lp.lfac_v[s] < 1.0f)
                        {
                            const float rq = base_v[s] * lp.g6_v[s];
                            if (r < rq)
                            {
                                const float c6s = pr.c6[s] * (1.0f / 6.0f), c12s = pr.c12[s] * (1.0f / 12.0f);
                                const float ri  = fep_rcp(rq);
                                const float ri3 = ri * ri * ri;
                                const float ri6 = ri3 * ri3;
                                const float ri7 = ri6 * ri;
                                const float ri8 = ri7 * ri;
                                const float t14 = c12s * ri7 * ri7 * r2;
                                const float t13 = c12s * ri7 * ri6 * r;
                                const float t12 = c12s * ri6 * ri6;
                                const float t8  = ri8 * c6s * r2;
                                const float t7  = ri7 * c6s * r;
                                const float t6  = ri6 * c6s;
                                const float quad = 156.0f * t14 - 42.0f * t8;
                                const float lin  = 168.0f * t13 - 48.0f * t7;
                                const float cst  = 91.0f * t12 - 28.0f * t6;
                                dv[s]            = 0.5f * quad - lin + cst;
                                xv2[s] = 28.0f * lp.gdl_v[s] * ((6.5f * t14 - t8) - (13.0f * t13 - 2.0f * t7) + (6.5f * t12 - t6));
                            }
                        }
                    }
                }
                acc[L::iVA + p] += dv[0];
                acc[L::iDV + p] += dv[1] - dv[0];
                acc[L::iCA + p] += dc[0];
                acc[L::iDCp + p] += dc[1] - dc[0];
                acc[L::iXC + p] += xc2[1] - xc2[0];
                acc[L::iXV + p] += xv2[1] - xv2[0];
            }
        }
    }

    /* per CTA: fp64 sums, then the lambda weights */
#pragma unroll
    for (int g = 0; g < L::N8; g++)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            const float w = fep_warp_sum(acc[8 * g + k]);
            if (lane == 0)
            {
                s_red[warp][8 * g + k] = w;
            }
        }
    }
    __syncthreads();
    if (tid < L::NACC)
    {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++)
        {
            s += (double)s_red[w][tid];
        }
        s_sum[tid] = s;
    }
    __syncthreads();
    if (tid < np)
    {
        const int          p  = tid;
        const LambdaPoint& lp = s_pts[p];
        const double CA = s_sum[L::iC] + s_sum[L::iCA + p], DC = s_sum[L::iDC] + s_sum[L::iDCp + p];
        const double GA = s_sum[L::iG] + s_sum[L::iVA + p], DG = s_sum[L::iDG] + s_sum[L::iDV + p];
        const double e  = (double)lp.lfac_c[0] * CA + (double)lp.lfac_c[1] * (CA + DC) + (double)lp.lfac_v[0] * GA
                         + (double)lp.lfac_v[1] * (GA + DG);
        const size_t o = (size_t)(3 * (p0 + p)) * gs.n_tiles + blockIdx.x;
        ka.for_part[o]                  = e;
        ka.for_part[o + gs.n_tiles]     = DC + s_sum[L::iXC + p];
        ka.for_part[o + 2 * gs.n_tiles] = DG + s_sum[L::iXV + p];
    }
    fep_pdl_wait();
}

template<bool EWALD, int C>
static void gs_launch_one(const KernelArgs& ka, const GapsysStep& gs, cudaStream_t stream, bool chained, int* occ)
{
    const size_t smem = fep_ring_bytes(FEP_GS_CTA / 32);
    if (occ)
    {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, fep_gapsys_foreign_kernel<EWALD, C>, FEP_GS_CTA, smem);
        return;
    }
    fep_launch_kernel_smem(fep_gapsys_foreign_kernel<EWALD, C>, dim3(gs.n_tiles, (ka.n_points + C - 1) / C), dim3(FEP_GS_CTA), smem,
                           stream, chained, ka, gs);
}

template<bool EWALD>
static bool gs_launch_size(const KernelArgs& ka, const GapsysStep& gs, int c, cudaStream_t stream, bool chained, int* occ)
{
    switch (c)
    {
        case 1: gs_launch_one<EWALD, 1>(ka, gs, stream, chained, occ); return true;
        case 2: gs_launch_one<EWALD, 2>(ka, gs, stream, chained, occ); return true;
        case 3: gs_launch_one<EWALD, 3>(ka, gs, stream, chained, occ); return true;
        case 4: gs_launch_one<EWALD, 4>(ka, gs, stream, chained, occ); return true;
        case 6: gs_launch_one<EWALD, 6>(ka, gs, stream, chained, occ); return true;
        case 8: gs_launch_one<EWALD, 8>(ka, gs, stream, chained, occ); return true;
        default: return false;
    }
}

/* points per launch for n_points lambda points: evenly sized chunks of at most FEP_GS_MAXC from the instantiated sizes */
extern "C" int fep_gapsys_chunk_size(int n_points)
{
    static const int sizes[] = { 1, 2, 3, 4, 6, 8 };
    const int        chunks  = (n_points + FEP_GS_MAXC - 1) / FEP_GS_MAXC;
    const int        need    = (n_points + chunks - 1) / std::max(chunks, 1);
    for (int c : sizes)
    {
        if (c >= need)
        {
            return c;
        }
    }
    return FEP_GS_MAXC;
}

extern "C" int fep_gapsys_ctas_per_sm(int elec_ewald, int c)
{
    KernelArgs ka{};
    GapsysStep gs{};
    int        occ = 0;
    const bool ok  = elec_ewald ? gs_launch_size<true>(ka, gs, c, nullptr, false, &occ) : gs_launch_size<false>(ka, gs, c, nullptr, false, &occ);
    return (ok && occ > 0) ? occ : 1;
}

/* all lambda points of a step in ONE launch: grid = ka.n_tiles CTAs x chunks of ka.chunk_points points (the points are
 * read from ka.pts on the device); pts = HOST copies of the points, for the candidate radii */
extern "C" int fep_launch_gapsys_foreign(const KernelArgs* kap, int elec_ewald, const LambdaPoint* pts, cudaStream_t stream,
                                         long long* counter, int chained)
{
    const KernelArgs& ka = *kap;
    GapsysStep        gs;
    gs.n_tiles = ka.n_tiles;
    for (int s = 0; s < 2; s++)
    {
        gs.g6c_max[s] = gs.g6v_max[s] = 0.0f;
        for (int p = 0; p < ka.n_points; p++)
        {
            gs.g6c_max[s] = std::max(gs.g6c_max[s], pts[p].g6_c[s]);
            gs.g6v_max[s] = std::max(gs.g6v_max[s], pts[p].g6_v[s]);
        }
    }
    const bool ok = elec_ewald ? gs_launch_size<true>(ka, gs, ka.chunk_points, stream, chained != 0, nullptr)
                               : gs_launch_size<false>(ka, gs, ka.chunk_points, stream, chained != 0, nullptr);
    if (!ok)
    {
        return (int)cudaErrorInvalidValue;
    }
    (*counter)++;
    return (int)cudaGetLastError();
}
