/*
 * fep_foreign_beutler.cu -- the energy-only foreign-lambda passes for the Beutler soft-core,
 * specialised so that the loop over lambda points is a handful of FP32/MUFU instructions.
 *
 * What is computed: freeenergydispatch.cpp:236-306 calling the energy-only flavour of
 * nb_free_energy_kernel (nb_free_energy.cpp:274-1187, computeForces == false) once per lambda
 * point.  How: everything that does not depend on lambda is evaluated once per pair and the
 * sums over pairs are kept PER STATE, because every lambda dependence outside the soft-core radius
 * is a weight that can be applied after the sum:
 *
 *   E(p)        = sum_s lfacC[s][p] * (C_s + Cp_s[p]) + lfacV[s][p] * (G_s + V_s[p])
 *   dVdl_coul(p) = (C_B + Cp_B[p]) - (C_A + Cp_A[p]),   dVdl_vdw(p) = (G_B + V_B[p]) - (G_A + V_A[p])
 *   (the kernel accumulates the state-A sums and the per-pair B-minus-A differences)
 *
 *   C_s    reaction-field / Ewald / exclusion terms that are linear in qq[s]   (:1023-1101)
 *          plus, when the Coulomb radius is not soft-cored (alphaCoul == 0), the whole
 *          Coulomb energy of state s
 *   G_s    LJ-PME grid correction, linear in c6grid[s]                         (:1103-1136)
 *   V_s[p] Lennard-Jones energy of state s with the soft-core radius of point p (:880-971)
 *   Cp_s[p] Coulomb energy of state s with the soft-core radius of point p     (:804-874), only
 *          when alphaCoul != 0
 *
 * (energy-only passes have no soft-core term in dV/dlambda, :1005-1013 with zero force terms).
 *
 * MODE 0: alphaCoul == 0 (the GROMACS default sc-coul = no): per point and state
 *         d = alphaVdwEff*sigma6*sclfacV + r^6 ; 1/d by MUFU.RCP ; LJ from 1/d : 8 instructions.
 * MODE 1: alphaCoul == alphaVdw and lambdaCoul == lambdaVdw at every point: one radius, the
 *         Coulomb part needs d^(-1/6) = ex2(-lg2(d)/6).
 * MODE 2: separate Coulomb and LJ radii.
 * The lambda factors of the chunk arrive as a __grid_constant__ kernel parameter, so they are
 * constant-bank operands of the FMAs (no loads in the inner loop); one launch per chunk of at
 * most FEP_FB_MAXC points.
 */
#include "fep_pair_math.cuh"

#define FULL_MASK 0xffffffffu
#define FEP_FB_MAXC 24

struct ForeignChunk
{
    float sclv[2][FEP_FB_MAXC]; /* soft-core lambda factor, vdw, per state */
    float sclc[2][FEP_FB_MAXC]; /* same for coulomb                         */
    float lfc[2][FEP_FB_MAXC];  /* {1-lambda_c, lambda_c}                   */
    float lfv[2][FEP_FB_MAXC];
    int   p0, np;               /* first point of the chunk, valid points   */
};

/* sums N8*8 per-lane values over the warp; afterwards lane l < 8 holds, for group g, the value
 * with index 8*g + 4*(l&1) + 2*((l>>1)&1) + ((l>>2)&1) */
template<int N8>
__device__ __forceinline__ void warp_sum_groups(float (&v)[N8 * 8], float (&out)[N8], int lane)
{
#pragma unroll
    for (int g = 0; g < N8; g++)
    {
        float a[4], b[2], c;
        {
            const bool up = lane & 1;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                const float send = up ? v[8 * g + i] : v[8 * g + i + 4];
                const float keep = up ? v[8 * g + i + 4] : v[8 * g + i];
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
        out[g] = c;
    }
}

/* lambda-independent data of one state of one pair */
struct StateConsts
{
    float c6_6, c12_12, shiftc, kv, kc, qe, qsh, qkrf;
};

/* One lambda point of one state: LJ energy vv (and Coulomb energy vc when the Coulomb radius is
 * soft-cored).  8 instructions in MODE 0. */
template<bool EWALD, int MODE>
__device__ __forceinline__ void fb_point(const StateConsts& st, float r6, float sclv, float sclc, float thr_v,
                                         float rcoulomb6, float& vv, float& vc)
{
    const float dv  = fmaf(st.kv, sclv, r6);
    const float ri6 = fminf(fep_rcp(dv), FEP_MAX_RINV6);
    vv              = fmaf(ri6, fmaf(st.c12_12, ri6, -st.c6_6), st.shiftc);
    vv              = dv < thr_v ? vv : 0.0f;
    if (MODE != 0)
    {
        const float dc  = (MODE == 1) ? dv : fmaf(st.kc, sclc, r6);
        const float lg  = fep_lg2(dc);
        const float ric = fep_ex2(lg * (-1.0f / 6.0f));
        if (EWALD)
        {
            vc = fmaf(st.qe, ric, st.qsh);
        }
        else
        {
            const float rc2 = fep_ex2(lg * (1.0f / 3.0f));
            vc              = fmaf(st.qe, ric, fmaf(st.qkrf, rc2, st.qsh));
            vc              = dc < rcoulomb6 ? vc : 0.0f;
        }
    }
}

template<bool EWALD, int MODE, int C>
__global__ void __launch_bounds__(FEP_FB_CTA, (((MODE == 0) ? 2 : 4) * C + 4 > 56) ? 2 : 4)
        fep_foreign_beutler_kernel(const __grid_constant__ KernelArgs ka, const __grid_constant__ ForeignChunk ch)
{
    constexpr int NPER = (MODE == 0) ? 2 : 4;      /* per-point accumulators: V_A DV (Cp_A DCp)   */
    constexpr int NACC = NPER * C + 4;             /* + C_A DC G_A DG                             */
    constexpr int N8   = (NACC + 7) / 8;
    constexpr int NW   = FEP_FB_CTA / 32;
    __shared__ float  s_red[NW][N8 * 8];
    __shared__ double s_sum[N8 * 8];

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    /* Sums over this thread's pairs.  State A sums and B-minus-A DIFFERENCES are accumulated (not
     * the two state sums): the difference is formed per pair, as in the reference (:1005-1020), so
     * pairs whose A and B parameters coincide cancel exactly instead of to rounding error.
     * Layout: [0,C) V_A, [C,2C) DV, MODE>0: [2C,3C) Cp_A, [3C,4C) DCp, then C_A DC G_A DG. */
    float acc[N8 * 8];
#pragma unroll
    for (int i = 0; i < N8 * 8; i++)
    {
        acc[i] = 0.0f;
    }
    constexpr int iCA = NPER * C, iDC = iCA + 1, iGA = iCA + 2, iDG = iCA + 3;

    const float thr_v = ka.vdw_ewald ? __int_as_float(0x7f800000) : ka.rvdw6; /* LJ-PME tests r, below */
    const int   base  = blockIdx.x * ka.tile_pairs;
    const int   end   = min(base + ka.tile_pairs, ka.n_pairs);

    for (int w0 = base + warp * 32; w0 < end; w0 += FEP_FB_CTA)
    {
        const int  slot   = w0 + lane;
        const bool active = slot < end;
        int        pj = 0, pe = 0;
        if (active)
        {
            pj = __ldg(ka.pair_j + slot);
            pe = __ldg(ka.pair_e + slot);
        }
        const bool   excluded = pj < 0;
        const int    cj       = pj & 0x7fffffff;
        const int4   en       = __ldg(ka.ent4 + pe);
        const float4 xi       = __ldg(ka.pos4 + en.x);
        const float4 sh       = ka.dyn->shiftvec[en.y];
        const float4 xj       = __ldg(ka.pos4 + cj);
        const float  dx = (sh.x + xi.x) - xj.x, dy = (sh.y + xi.y) - xj.y, dz = (sh.z + xi.z) - xj.z;
        float        r2       = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const bool   within   = r2 < ka.rcut_max2;
        const bool   contrib  = active && (within || excluded); /* :667 */
        if (!__any_sync(FULL_MASK, contrib))
        {
            continue;
        }
        const float4 pi = __ldg(ka.par4 + en.x);
        const float4 pq = __ldg(ka.par4 + cj);
        const float4 ta = __ldg(ka.typetab + (ka.ntype * __float_as_int(pi.z) + __float_as_int(pq.z)));
        const float4 tb = __ldg(ka.typetab + (ka.ntype * __float_as_int(pi.w) + __float_as_int(pq.w)));
        const float  m  = contrib ? 1.0f : 0.0f;
        const float  qq[2]   = { (ka.epsfac * pi.x) * pq.x * m, (ka.epsfac * pi.y) * pq.y * m };
        const float  c6[2]   = { ta.x, tb.x }, c12[2] = { ta.y, tb.y }, sig6[2] = { ta.z, tb.z };
        const float  c6g[2]  = { ta.w * m, tb.w * m };
        const bool   hard    = (ta.y > 0.0f && tb.y > 0.0f); /* :597-628 */
        const float  a_v     = hard ? 0.0f : ka.alpha_v;
        const float  a_c     = hard ? 0.0f : ka.alpha_c;
        const bool   self    = (en.x == cj);

        FepPair pr;
        r2      = fmaxf(r2, FEP_MIN_RSQ);
        pr.r2   = r2;
        pr.rinv = fep_rsqrt(r2);
        pr.r    = r2 * pr.rinv;
        const float r6   = r2 * r2 * r2;
        const bool  incl = contrib && within && !excluded;

        /* lambda-independent correction terms, linear in qq[s] / c6grid[s] */
        {
            float xc, fc, xv, fv;
            fep_corrections<EWALD, false>(ka, pr, excluded, self, xc, fc, xv, fv);
            const float cA = qq[0] * xc, gA = c6g[0] * xv;
            acc[iCA] += cA;
            acc[iDC] += fmaf(qq[1], xc, -cA);
            acc[iGA] += gA;
            acc[iDG] += fmaf(c6g[1], xv, -gA);
        }

        StateConsts st[2];
        bool        elec_on[2], vdw_on[2];
        float       vc0[2];
#pragma unroll
        for (int s = 0; s < 2; s++)
        {
            const bool nz = incl && (qq[s] != 0.0f || c6[s] != 0.0f || c12[s] != 0.0f); /* :747-752 */
            /* lambda-independent parts of the interaction tests (:805-812, :880-890) */
            elec_on[s] = nz && qq[s] != 0.0f;
            vdw_on[s]  = nz && (c6[s] != 0.0f || c12[s] != 0.0f);
            if (EWALD || MODE == 0)
            {
                elec_on[s] = elec_on[s] && pr.r < ka.rcoulomb;
            }
            if (ka.vdw_ewald)
            {
                vdw_on[s] = vdw_on[s] && pr.r < ka.rvdw;
            }
            st[s].qe     = elec_on[s] ? qq[s] : 0.0f;
            st[s].c6_6   = vdw_on[s] ? c6[s] * (1.0f / 6.0f) : 0.0f;
            st[s].c12_12 = vdw_on[s] ? c12[s] * (1.0f / 12.0f) : 0.0f;
            st[s].shiftc = st[s].c12_12 * ka.rep_cpot - st[s].c6_6 * ka.disp_cpot;
            if (ka.vdw_ewald)
            {
                st[s].shiftc = fmaf(vdw_on[s] ? c6g[s] : 0.0f, ka.sh_lj_ewald * (1.0f / 6.0f), st[s].shiftc);
            }
            st[s].kv   = a_v * sig6[s];
            st[s].kc   = a_c * sig6[s];
            st[s].qsh  = EWALD ? -st[s].qe * ka.sh_ewald : -st[s].qe * ka.crf;
            st[s].qkrf = st[s].qe * ka.krf;
            /* Coulomb radius not soft-cored: rC == r, the whole term is lambda-independent */
            vc0[s] = EWALD ? st[s].qe * (pr.rinv - ka.sh_ewald) : st[s].qe * (pr.rinv + fmaf(ka.krf, r2, -ka.crf));
        }
        if (MODE == 0)
        {
            acc[iCA] += vc0[0];
            acc[iDC] += vc0[1] - vc0[0];
        }
        /* a state nobody in the warp needs is skipped; the choice is made once per 32 pairs */
        const bool needA = __any_sync(FULL_MASK, vdw_on[0] || (MODE != 0 && elec_on[0]));
        const bool needB = __any_sync(FULL_MASK, vdw_on[1] || (MODE != 0 && elec_on[1]));
        if (needA && needB)
        {
#pragma unroll
            for (int p = 0; p < C; p++)
            {
                float vvA, vvB, vcA = 0.0f, vcB = 0.0f;
                fb_point<EWALD, MODE>(st[0], r6, ch.sclv[0][p], ch.sclc[0][p], thr_v, ka.rcoulomb6, vvA, vcA);
                fb_point<EWALD, MODE>(st[1], r6, ch.sclv[1][p], ch.sclc[1][p], thr_v, ka.rcoulomb6, vvB, vcB);
                acc[p]     += vvA;
                acc[C + p] += vvB - vvA;
                if (MODE != 0)
                {
                    acc[2 * C + p] += vcA;
                    acc[3 * C + p] += vcB - vcA;
                }
            }
        }
        else if (needA)
        {
#pragma unroll
            for (int p = 0; p < C; p++)
            {
                float vvA, vcA = 0.0f;
                fb_point<EWALD, MODE>(st[0], r6, ch.sclv[0][p], ch.sclc[0][p], thr_v, ka.rcoulomb6, vvA, vcA);
                acc[p]     += vvA;
                acc[C + p] -= vvA;
                if (MODE != 0)
                {
                    acc[2 * C + p] += vcA;
                    acc[3 * C + p] -= vcA;
                }
            }
        }
        else if (needB)
        {
#pragma unroll
            for (int p = 0; p < C; p++)
            {
                float vvB, vcB = 0.0f;
                fb_point<EWALD, MODE>(st[1], r6, ch.sclv[1][p], ch.sclc[1][p], thr_v, ka.rcoulomb6, vvB, vcB);
                acc[C + p] += vvB;
                if (MODE != 0)
                {
                    acc[3 * C + p] += vcB;
                }
            }
        }
    }

    float red[N8];
    warp_sum_groups<N8>(acc, red, lane);
    if (lane < 8)
    {
        const int k = 4 * (lane & 1) + 2 * ((lane >> 1) & 1) + ((lane >> 2) & 1);
#pragma unroll
        for (int g = 0; g < N8; g++)
        {
            s_red[warp][8 * g + k] = red[g];
        }
    }
    __syncthreads();
    if (tid < NACC)
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
    if (tid < ch.np)
    {
        const int    p  = tid;
        const double CA = s_sum[iCA] + (MODE != 0 ? s_sum[2 * C + p] : 0.0);
        const double DC = s_sum[iDC] + (MODE != 0 ? s_sum[3 * C + p] : 0.0);
        const double GA = s_sum[iGA] + s_sum[p];
        const double DG = s_sum[iDG] + s_sum[C + p];
        /* E = lfacC[A] C_A + lfacC[B] C_B + lfacV[A] G_A + lfacV[B] G_B with X_B = X_A + DX */
        const double e  = (double)ch.lfc[0][p] * CA + (double)ch.lfc[1][p] * (CA + DC) + (double)ch.lfv[0][p] * GA
                         + (double)ch.lfv[1][p] * (GA + DG);
        const size_t o  = (size_t)(3 * (ch.p0 + p)) * ka.n_tiles + blockIdx.x;
        ka.for_part[o]                  = e;
        ka.for_part[o + ka.n_tiles]     = DC;
        ka.for_part[o + 2 * ka.n_tiles] = DG;
    }
}

/* ------------------------------------------------------------------------------------------- */
template<bool EWALD, int MODE, int C>
static void launch_one(const KernelArgs& ka, const ForeignChunk& ch, cudaStream_t stream)
{
    fep_foreign_beutler_kernel<EWALD, MODE, C><<<ka.n_tiles, FEP_FB_CTA, 0, stream>>>(ka, ch);
}

template<bool EWALD, int MODE>
static bool launch_size(const KernelArgs& ka, const ForeignChunk& ch, int c, cudaStream_t stream)
{
    switch (c)
    {
#define FEP_FB_CASE(N) \
    case N: launch_one<EWALD, MODE, N>(ka, ch, stream); return true;
        FEP_FB_CASE(1)
        FEP_FB_CASE(2)
        FEP_FB_CASE(3)
        FEP_FB_CASE(4)
        FEP_FB_CASE(6)
        FEP_FB_CASE(7)
        FEP_FB_CASE(8)
        FEP_FB_CASE(11)
        FEP_FB_CASE(14)
        FEP_FB_CASE(16)
        FEP_FB_CASE(21)
        FEP_FB_CASE(24)
#undef FEP_FB_CASE
        default: return false;
    }
}

static const int c_sizes[] = { 1, 2, 3, 4, 6, 7, 8, 11, 14, 16, 21, 24 };

extern "C" int fep_foreign_beutler_chunk_size(int n_points, int n_chunks_wanted)
{
    /* smallest supported chunk size that covers n_points with at most... the wanted chunk count */
    if (n_chunks_wanted < 1)
    {
        n_chunks_wanted = 1;
    }
    const int need = (n_points + n_chunks_wanted - 1) / n_chunks_wanted;
    for (int c : c_sizes)
    {
        if (c >= need)
        {
            return c;
        }
    }
    return FEP_FB_MAXC;
}

/* Launches the foreign passes for all points; `pts` are the HOST copies of the lambda points.
 * Returns cudaSuccess (0), or -1 when this specialisation does not cover the case. */
extern "C" int fep_launch_foreign_beutler(const KernelArgs* kap, int elec_ewald, int mode, const LambdaPoint* pts,
                                          cudaStream_t stream, long long* counter)
{
    const KernelArgs& ka = *kap;
    const int         c  = ka.chunk_points;
    for (int p0 = 0; p0 < ka.n_points; p0 += c)
    {
        ForeignChunk ch;
        ch.p0 = p0;
        ch.np = (ka.n_points - p0 < c) ? ka.n_points - p0 : c;
        for (int p = 0; p < FEP_FB_MAXC; p++)
        {
            /* padding points repeat the last valid one; their results are not written */
            int q = p0 + (p < ch.np ? p : ch.np - 1);
            for (int s = 0; s < 2; s++)
            {
                ch.sclv[s][p] = pts[q].sclfac_v[s];
                ch.sclc[s][p] = pts[q].sclfac_c[s];
                ch.lfc[s][p]  = pts[q].lfac_c[s];
                ch.lfv[s][p]  = pts[q].lfac_v[s];
            }
        }
        bool ok;
        if (elec_ewald)
        {
            ok = mode == 0   ? launch_size<true, 0>(ka, ch, c, stream)
                 : mode == 1 ? launch_size<true, 1>(ka, ch, c, stream)
                             : launch_size<true, 2>(ka, ch, c, stream);
        }
        else
        {
            ok = mode == 0   ? launch_size<false, 0>(ka, ch, c, stream)
                 : mode == 1 ? launch_size<false, 1>(ka, ch, c, stream)
                             : launch_size<false, 2>(ka, ch, c, stream);
        }
        if (!ok)
        {
            return -1;
        }
        (*counter)++;
    }
    return (int)cudaGetLastError();
}
