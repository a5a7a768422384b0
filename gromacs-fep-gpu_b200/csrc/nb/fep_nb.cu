/*
 * fep_nb.cu -- the non-perturbed cluster-pair kernel next to the FEP path and its C-ABI
 * (include/fepb200_nb.h), hand-written for sm_100a.
 *
 * WHAT is computed is the reference's kernel for GPU-layout pair lists,
 *   nbnxn_kernel_gpu_ref()        src/gromacs/nbnxm/kernels_reference/kernel_gpu_ref.cpp:54-354
 * on atoms whose perturbed members were masked like
 *   nbnxn_atomdata_mask_fep()     src/gromacs/nbnxm/atomdata.cpp:930-964
 * (its CUDA twin in the fork: nbnxm/cuda/nbnxm_cuda_kernel.cuh, launched at nbnxm_cuda.cu:642-871).
 *
 * HOW is ours.  One WARP owns a work item = one i-super-cluster entry x a chunk of its packed j-cluster entries
 * (items are cut at set_pairlist so that there are several per resident warp: no host-side list splitting needed,
 * no CTA-wide synchronisation anywhere).  A lane is (ii = lane%8, jq = lane/8): it keeps i atom ii of all 8 i-clusters
 * of the super-cluster -- coordinates, charge, type row and force accumulators -- in REGISTERS for the whole item, and
 * for each listed j-cluster the two j atoms jq and jq+4, i.e. the two halves of the reference's cluster-pair split:
 * lane == word index of nbnxn_excl_t::pair in both halves, so a packed entry's interaction bits are two coalesced loads.
 * Per (i-cluster, j-cluster) pair that the imask lists (a warp-uniform test: unlisted cluster pairs cost nothing) a
 * lane evaluates two atom pairs with straight-line code (cut-offs, exclusion bit and the diagonal rule are factors and
 * selects, not branches).  j forces: 6 values per lane are reduced over the 8 lanes that share the j atoms with a
 * 6-shuffle reduce-scatter that leaves each value on its own lane -> one atomic add per value.  i forces stay in
 * registers until the item ends: an 18-shuffle reduce-scatter over the 4 lanes that share an i atom, 6 atomic adds per lane.
 * Ewald real space is analytical: the rational fit B(w) of fep_pair_math.cuh (tools/fit_ewald_rational.py) for the force,
 * V(w) for the energy -- no table, no texture.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/fepb200_nb.h"

namespace
{

constexpr int   CL            = FEPB200_NB_CLUSTER_SIZE;
constexpr int   NCL           = FEPB200_NB_CLUSTERS_PER_SUPER;
constexpr int   CENTRAL_SHIFT = 22;
constexpr int   WARPS_PER_CTA = 4;
constexpr float MIN_RSQ       = FEPB200_NB_MIN_RSQ;

struct NbItem
{
    int entry;    /* index into sci[] */
    int cj_begin; /* packed j-cluster entries [cj_begin, cj_end) */
    int cj_end;
    int self;     /* 1: this item books the charge self term of its super-cluster (kernel_gpu_ref.cpp:122-147) */
};

struct NbConsts
{
    float epsfac, rc2, rv2;
    float k_rf, c_rf;
    float sh_ewald, beta, beta2, beta3;
    float disp_cpot, rep_cpot;
    float self_coef; /* -epsfac beta / sqrt(pi)  or  -epsfac c_rf / 2 */
    /* Lennard-Jones modifiers of the reference's CUDA kernels (nbnxm_cuda_kernel_utils.cuh:104-211): force switch
     * (dispersion / repulsion c2, c3 of mdtypes/interaction_const.cpp:216-230) and potential switch (c3, c4, c5 of :232-245) */
    float rvdw_switch, dsp_c2, dsp_c3, rep_c2, rep_c3, sw_c3, sw_c4, sw_c5;
    int   ntype;
};

struct NbArgs
{
    const float4* __restrict__ xq;
    const int* __restrict__ type;
    const float2* __restrict__ nbfp;
    const float* __restrict__ shiftvec;
    const fepb200_nb_sci* __restrict__ sci;
    const int4* __restrict__ cj; /* two int4 per fepb200_nb_cj_packed */
    const fepb200_nb_excl* __restrict__ excl;
    const NbItem* __restrict__ items;
    int      nitems;
    unsigned int* next_item; /* work queue head of THIS launch (zero at its start) */
    unsigned int* next_reset; /* the head the next launch on this stream will use: zeroed by this one */
    float*        e_el_f;    /* float accumulators instead of `energies` (NBAtomDataGpu::eElec / eLJ), or NULL */
    float*        e_lj_f;
    float*   f;      /* float3[natoms], added into */
    float*   fshift; /* float[135] or NULL */
    double*  energies; /* {vc, vvdw} or NULL */
    NbConsts c;
};

__device__ __forceinline__ float nb_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float nb_rsqrt(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

/* Ewald real space, w = beta^2 r^2 <= 16: V(w) = erf(z)/z and B(w) = 2 dV/dw as rational functions -- the same fits
 * (our own, tools/fit_ewald_rational.py) the perturbed-pair kernels use, fep_pair_math.cuh:66-112. */
__device__ __forceinline__ float nb_ewald_V(float w)
{
    float p = 1.914866538e-08f;
    p       = fmaf(p, w, -1.938895923e-06f);
    p       = fmaf(p, w, 1.461872746e-04f);
    p       = fmaf(p, w, 3.943561305e-03f);
    p       = fmaf(p, w, 5.137383335e-02f);
    p       = fmaf(p, w, 2.060183881e-01f);
    p       = fmaf(p, w, 1.128379099e+00f);
    float q = 1.009842853e-03f;
    q       = fmaf(q, w, 1.486172240e-02f);
    q       = fmaf(q, w, 1.175093691e-01f);
    q       = fmaf(q, w, 5.159104191e-01f);
    q       = fmaf(q, w, 1.0f);
    return p * nb_rcp(q);
}
__device__ __forceinline__ float nb_ewald_B(float w)
{
    float p = -1.081098709e-08f;
    p       = fmaf(p, w, 1.028332784e-06f);
    p       = fmaf(p, w, -5.036981975e-05f);
    p       = fmaf(p, w, 2.020152613e-04f);
    p       = fmaf(p, w, -1.807793702e-02f);
    p       = fmaf(p, w, 3.588346990e-02f);
    p       = fmaf(p, w, -7.522528250e-01f);
    float q = 1.302164612e-04f;
    q       = fmaf(q, w, 2.096415634e-03f);
    q       = fmaf(q, w, 2.163110569e-02f);
    q       = fmaf(q, w, 1.411150645e-01f);
    q       = fmaf(q, w, 5.523007151e-01f);
    q       = fmaf(q, w, 1.0f);
    return p * nb_rcp(q);
}

#ifndef NB_ROWSKIP
#define NB_ROWSKIP 1
#endif
#define FULL 0xffffffffu

struct F3
{
    float x, y, z;
};

/* One atom pair.  `bit`: the pair interacts (false: excluded pair, which still gets the reaction-field / Ewald
 * correction); `skip`: lower triangle of a cluster against itself in the central cell.  Everything a pair can be ruled out by
 * is a select on the result, never a factor: a pair far outside the cut-off (filler atoms) may produce inf / NaN on the way.
 * SAMECUT: rvdw == rcoulomb, one test for both.  `nb`: the {6 C6, 12 C12} table, in shared memory when TABSMEM. */
template<bool EWALD, bool ENERGY, bool SAMECUT, bool TABSMEM, int VDWMOD>
__device__ __forceinline__ void nb_pair(const NbConsts& c, const float2* __restrict__ nb, float xi, float yi, float zi,
                                        float qi, int ti, const float4& xj, int tj, bool bit, bool skip, F3& fi, F3& fj,
                                        float& e_el, float& e_lj)
{
    const float  dx = xi - xj.x, dy = yi - xj.y, dz = zi - xj.z;
    const float  r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    const bool   in = (r2 < c.rc2) && !skip;
#if NB_ROWSKIP
    /* the 32 atom pairs of this warp instruction (8 i atoms x 4 j atoms) are often all outside the cut-off -- a list is
     * made with a buffer around it: then the whole evaluation is skipped (one vote, one uniform branch) */
    if (!__any_sync(FULL, in))
    {
        return;
    }
#endif
    const float  r2c   = fmaxf(r2, MIN_RSQ);
    const float  rinv  = nb_rsqrt(r2c);
    const float  rinv2 = rinv * rinv;
    const float  qq    = qi * xj.w;
    const float2 cc    = TABSMEM ? nb[ti + tj] : __ldg(nb + ti + tj);
    const float  brinv = bit ? rinv : 0.0f;
    float        fs, v_el;
    if (EWALD)
    {
        const float w = c.beta2 * r2c;
        /* force factor of erf(beta r)/r: f_lr = -beta^3 B(w) */
        fs = qq * fmaf(brinv, rinv2, c.beta3 * nb_ewald_B(w));
        if (ENERGY)
        {
            v_el = qq * (brinv - c.beta * nb_ewald_V(w) - (bit ? c.sh_ewald : 0.0f));
        }
    }
    else
    {
        const float kr2 = c.k_rf * r2c;
        fs              = qq * fmaf(-2.0f, kr2, brinv) * rinv2;
        if (ENERGY)
        {
            v_el = qq * (brinv + kr2 - c.c_rf);
        }
    }
    const bool  inlj = SAMECUT ? in : (in && (r2c < c.rv2));
    const float r6   = bit ? rinv2 * rinv2 * rinv2 : 0.0f;
    float       v6 = 0.0f, v12 = 0.0f, flj, vlj = 0.0f;
    if (ENERGY || VDWMOD == 2)
    {
        v6  = cc.x * r6;
        v12 = cc.y * r6 * r6;
        flj = (v12 - v6) * rinv2;
        vlj = fmaf(bit ? cc.y : 0.0f, c.rep_cpot, v12) * (1.0f / 12.0f) - fmaf(bit ? cc.x : 0.0f, c.disp_cpot, v6) * (1.0f / 6.0f);
    }
    else
    {
        flj = fmaf(cc.y, r6, -cc.x) * (r6 * rinv2); /* (12 C12 r^-12 - 6 C6 r^-6) r^-2 */
    }
    if (VDWMOD != 0)
    {
        /* like the reference's CUDA kernels: cc = {6 C6, 12 C12} are their c6 / c12, the switch terms do not carry the
         * interaction bit (an excluded pair in the switching region has zero parameters or is a bonded neighbour) */
        const float sd = fmaxf(fmaf(r2c, rinv, -c.rvdw_switch), 0.0f);
        if (VDWMOD == 1)
        {
            flj += (cc.y * fmaf(c.rep_c3, sd, c.rep_c2) - cc.x * fmaf(c.dsp_c3, sd, c.dsp_c2)) * (sd * sd * rinv);
            if (ENERGY)
            {
                vlj += (cc.x * fmaf(c.dsp_c3 * 0.25f, sd, c.dsp_c2 * (1.0f / 3.0f))
                        - cc.y * fmaf(c.rep_c3 * 0.25f, sd, c.rep_c2 * (1.0f / 3.0f)))
                       * (sd * sd * sd);
            }
        }
        else
        {
            const float sw  = fmaf(fmaf(fmaf(c.sw_c5, sd, c.sw_c4), sd, c.sw_c3), sd * sd * sd, 1.0f);
            const float dsw = fmaf(fmaf(5.0f * c.sw_c5, sd, 4.0f * c.sw_c4), sd, 3.0f * c.sw_c3) * (sd * sd);
            flj             = fmaf(flj, sw, -rinv * vlj * dsw);
            vlj *= sw;
        }
    }
    if (SAMECUT)
    {
        fs = in ? fs + flj : 0.0f;
    }
    else
    {
        fs += inlj ? flj : 0.0f;
        fs = in ? fs : 0.0f;
    }
    if (ENERGY)
    {
        /* the reference books the Coulomb energy inside its `rsq < rvdw2` branch (kernel_gpu_ref.cpp:264-287) */
        e_el += inlj ? v_el : 0.0f;
        e_lj += inlj ? vlj : 0.0f;
    }
    fi.x = fmaf(fs, dx, fi.x);
    fi.y = fmaf(fs, dy, fi.y);
    fi.z = fmaf(fs, dz, fi.z);
    fj.x = fmaf(-fs, dx, fj.x);
    fj.y = fmaf(-fs, dy, fj.y);
    fj.z = fmaf(-fs, dz, fj.z);
}


constexpr int ATOMS_SC = CL * NCL; /* 64 atoms per super-cluster */

/* Four CTAs per SM: with 128 registers every per-lane constant stays in a register; a build for five (96 registers)
 * recomputes some of them in every i-cluster block and was 19 % slower (profiles/r02_nb_kernel_variants.txt).
 * VDWMOD: 0 plain / potential shift (what nbnxn_kernel_gpu_ref computes), 1 force switch, 2 potential switch. */
template<bool EWALD, bool ENERGY, bool SAMECUT, bool TABSMEM, int VDWMOD>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4) fep_nb_kernel(const NbArgs a)
{
    /* per warp: the 64 i atoms of its item (shifted x, y, z and epsfac q; type row offset); then the LJ table */
    extern __shared__ float4 s_dyn[];
    int*    s_ti_all = reinterpret_cast<int*>(s_dyn + WARPS_PER_CTA * ATOMS_SC);
    float2* s_nbfp   = reinterpret_cast<float2*>(s_ti_all + WARPS_PER_CTA * ATOMS_SC);
    if (TABSMEM)
    {
        for (int k = threadIdx.x; k < a.c.ntype * a.c.ntype; k += blockDim.x)
        {
            s_nbfp[k] = a.nbfp[k];
        }
        __syncthreads();
    }
    const float2* __restrict__ nbtab = TABSMEM ? s_nbfp : a.nbfp;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        /* two queue heads take turns: launches of a handle are ordered on its stream, so the next launch finds its head at
         * zero without a memset in between */
        *a.next_reset = 0u;
    }
    const int lane   = threadIdx.x & 31;
    const int ii     = lane & 7;
    const int jq     = lane >> 3;
    float4*   s_xi   = s_dyn + (threadIdx.x >> 5) * ATOMS_SC;
    int*      s_ti   = s_ti_all + (threadIdx.x >> 5) * ATOMS_SC;
    const NbConsts& c = a.c;

    /* items differ in cost (how many of the 8 x 4 cluster pairs of a packed entry are listed): every warp takes the next
     * item from one queue, so that all SMs finish together */
    for (;;)
    {
        int item = 0;
        if (lane == 0)
        {
            item = (int)atomicAdd(a.next_item, 1u);
        }
        item = __shfl_sync(FULL, item, 0);
        if (item >= a.nitems)
        {
            break;
        }
        const NbItem         it    = a.items[item];
        const fepb200_nb_sci e     = a.sci[it.entry];
        const float          shx   = __ldg(a.shiftvec + 3 * e.shift);
        const float          shy   = __ldg(a.shiftvec + 3 * e.shift + 1);
        const float          shz   = __ldg(a.shiftvec + 3 * e.shift + 2);
        const int            ci0   = e.sci * NCL;
        const bool           centr = e.shift == CENTRAL_SHIFT;

        /* the first packed entry and the atoms of its first j-cluster are on their way while the i atoms are staged */
        int          g    = it.cj_begin;
        int4         cjv  = __ldg(a.cj + 2 * g);
        int4         ime  = __ldg(a.cj + 2 * g + 1);
        unsigned int wex0 = __ldg(&a.excl[ime.y].pair[lane]);
        unsigned int wex1 = __ldg(&a.excl[ime.w].pair[lane]);
        int          cjn_n = cjv.x;
        float4       xa_n  = __ldg(a.xq + cjn_n * CL + jq);
        float4       xb_n  = __ldg(a.xq + cjn_n * CL + jq + CL / 2);
        int          ta_n  = __ldg(a.type + cjn_n * CL + jq);
        int          tb_n  = __ldg(a.type + cjn_n * CL + jq + CL / 2);

        /* i atoms: two per lane, coalesced, into this warp's shared-memory block */
        __syncwarp();
        float q2 = 0.0f;
#pragma unroll
        for (int k = 0; k < ATOMS_SC / 32; k++)
        {
            const int ia = ci0 * CL + lane + 32 * k;
            float4    v  = __ldg(a.xq + ia);
            q2           = fmaf(v.w, v.w, q2);
            v.x += shx;
            v.y += shy;
            v.z += shz;
            v.w *= c.epsfac;
            s_xi[lane + 32 * k] = v;
            s_ti[lane + 32 * k] = c.ntype * __ldg(a.type + ia);
        }
        __syncwarp();
        F3 fi[NCL];
#pragma unroll
        for (int im = 0; im < NCL; im++)
        {
            fi[im].x = fi[im].y = fi[im].z = 0.0f;
        }
        float e_el = 0.0f, e_lj = 0.0f;
        if (ENERGY && it.self)
        {
            e_el = c.self_coef * q2; /* every lane holds two of the 64 atoms */
        }

        for (; g < it.cj_end; g++)
        {
            /* the next packed entry is fetched a whole entry ahead, its exclusion words half an entry ahead */
            const bool   more  = g + 1 < it.cj_end;
            int4         cjv_n = cjv, ime_n = ime;
            unsigned int wex0_n = wex0, wex1_n = wex1;
            if (more)
            {
                cjv_n = __ldg(a.cj + 2 * (g + 1));
                ime_n = __ldg(a.cj + 2 * (g + 1) + 1);
            }
            /* the two halves of a cluster pair carry their own mask once the fork's pruning kernels have been at the list
             * (each of its warps prunes its half): this kernel does both halves in one warp and needs their union */
            const unsigned int imask = (unsigned int)ime.x | (unsigned int)ime.z;
            /* NOT unrolled: the body below (8 i-clusters x 2 atom pairs, straight line) is 14 KB of code; four copies
             * of it do not fit the 32 KB instruction cache level and the warps of an SM then wait for fetches */
#pragma unroll 1
            for (int jm = 0; jm < FEPB200_NB_JGROUP_SIZE; jm++)
            {
                /* this j-cluster's two atoms per lane were loaded one j-cluster ago; now the next one's */
                const int    cjn = cjn_n;
                const float4 xa = xa_n, xb = xb_n;
                const int    ta = ta_n, tb = tb_n;
                const int    ja = cjn * CL + jq;
                if (jm < 3 || more)
                {
                    cjn_n = jm == 0 ? cjv.y : (jm == 1 ? cjv.z : (jm == 2 ? cjv.w : cjv_n.x));
                    xa_n  = __ldg(a.xq + cjn_n * CL + jq);
                    xb_n  = __ldg(a.xq + cjn_n * CL + jq + CL / 2);
                    ta_n  = __ldg(a.type + cjn_n * CL + jq);
                    tb_n  = __ldg(a.type + cjn_n * CL + jq + CL / 2);
                }
                if (jm == 2 && more)
                {
                    wex0_n = __ldg(&a.excl[ime_n.y].pair[lane]);
                    wex1_n = __ldg(&a.excl[ime_n.w].pair[lane]);
                }
                const unsigned int m8 = (imask >> (NCL * jm)) & 0xffu;
                if (m8 == 0)
                {
                    continue;
                }
                F3           fa = { 0.0f, 0.0f, 0.0f }, fb = { 0.0f, 0.0f, 0.0f };
                const unsigned int ea = wex0 >> (NCL * jm), eb = wex1 >> (NCL * jm);
                /* a cluster against itself in the central cell (at most one i-cluster of the eight): only the upper
                 * triangle of its atom pairs exists (kernel_gpu_ref.cpp:196-199).  One bit mask per lane and j atom,
                 * tested with the i-cluster's bit like the exclusion words. */
                const int          dself = centr ? cjn - ci0 : NCL;
                const unsigned int sbit  = (dself >= 0 && dself < NCL) ? (1u << dself) : 0u;
                const unsigned int ska = (jq <= ii) ? sbit : 0u, skb = (jq + CL / 2 <= ii) ? sbit : 0u;
#pragma unroll
                for (int im = 0; im < NCL; im++)
                {
                    if (m8 & (1u << im))
                    {
                        const float4 vi = s_xi[im * CL + ii];
                        const int    ti = s_ti[im * CL + ii];
                        nb_pair<EWALD, ENERGY, SAMECUT, TABSMEM, VDWMOD>(c, nbtab, vi.x, vi.y, vi.z, vi.w, ti, xa, ta,
                                                                 (ea >> im) & 1u, (ska >> im) & 1u, fi[im], fa, e_el, e_lj);
                        nb_pair<EWALD, ENERGY, SAMECUT, TABSMEM, VDWMOD>(c, nbtab, vi.x, vi.y, vi.z, vi.w, ti, xb, tb,
                                                                 (eb >> im) & 1u, (skb >> im) & 1u, fi[im], fb, e_el, e_lj);
                    }
                }
                /* j forces: reduce {fa, fb} over the 8 lanes ii = 0..7 that hold the same two j atoms, leaving each of
                 * the 6 sums on its own lane.  xor 4: atom a stays with ii < 4, atom b with ii >= 4. */
                const bool hi = ii & 4;
                float v0 = (hi ? fb.x : fa.x) + __shfl_xor_sync(FULL, hi ? fa.x : fb.x, 4);
                float v1 = (hi ? fb.y : fa.y) + __shfl_xor_sync(FULL, hi ? fa.y : fb.y, 4);
                float v2 = (hi ? fb.z : fa.z) + __shfl_xor_sync(FULL, hi ? fa.z : fb.z, 4);
                /* xor 2: (x, y) stay with bit 1 clear, z goes to bit 1 set */
                const bool  m  = ii & 2;
                const float s1 = __shfl_xor_sync(FULL, m ? v0 : v2, 2);
                const float s2 = __shfl_xor_sync(FULL, v1, 2);
                const float w0v = (m ? v2 : v0) + s1;
                const float w1v = v1 + s2;
                /* xor 1: x on even, y on odd lanes; z complete on both lanes of its pair */
                const bool  o   = ii & 1;
                const float r   = __shfl_xor_sync(FULL, m ? w0v : (o ? w0v : w1v), 1);
                const float sum = (m ? w0v : (o ? w1v : w0v)) + r;
                if (!(m && o))
                {
                    const int atom = ja + (hi ? CL / 2 : 0);
                    const int comp = m ? 2 : (o ? 1 : 0);
                    atomicAdd(a.f + 3 * (size_t)atom + comp, sum);
                }
            }
            cjv  = cjv_n;
            ime  = ime_n;
            wex0 = wex0_n;
            wex1 = wex1_n;
        }

        /* shift force of the item: everything its i atoms received */
        if (a.fshift != nullptr)
        {
            float sx = 0.0f, sy = 0.0f, sz = 0.0f;
#pragma unroll
            for (int im = 0; im < NCL; im++)
            {
                sx += fi[im].x;
                sy += fi[im].y;
                sz += fi[im].z;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                sx += __shfl_xor_sync(FULL, sx, d);
                sy += __shfl_xor_sync(FULL, sy, d);
                sz += __shfl_xor_sync(FULL, sz, d);
            }
            if (lane < 3)
            {
                atomicAdd(a.fshift + 3 * e.shift + lane, lane == 0 ? sx : (lane == 1 ? sy : sz));
            }
        }
        /* i forces: the 4 lanes jq = 0..3 hold partial sums for the same 8 i atoms (one per i-cluster): reduce-scatter,
         * xor 16 keeps i-clusters 0..3 on jq < 2 and 4..7 on jq >= 2, xor 8 then halves again: 2 i-clusters per lane */
        {
            const bool h16 = lane & 16;
            F3         g4[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                g4[k].x = (h16 ? fi[k + 4].x : fi[k].x) + __shfl_xor_sync(FULL, h16 ? fi[k].x : fi[k + 4].x, 16);
                g4[k].y = (h16 ? fi[k + 4].y : fi[k].y) + __shfl_xor_sync(FULL, h16 ? fi[k].y : fi[k + 4].y, 16);
                g4[k].z = (h16 ? fi[k + 4].z : fi[k].z) + __shfl_xor_sync(FULL, h16 ? fi[k].z : fi[k + 4].z, 16);
            }
            const bool h8 = lane & 8;
#pragma unroll
            for (int k = 0; k < 2; k++)
            {
                const float fx = (h8 ? g4[k + 2].x : g4[k].x) + __shfl_xor_sync(FULL, h8 ? g4[k].x : g4[k + 2].x, 8);
                const float fy = (h8 ? g4[k + 2].y : g4[k].y) + __shfl_xor_sync(FULL, h8 ? g4[k].y : g4[k + 2].y, 8);
                const float fz = (h8 ? g4[k + 2].z : g4[k].z) + __shfl_xor_sync(FULL, h8 ? g4[k].z : g4[k + 2].z, 8);
                const int   im = (h16 ? 4 : 0) + (h8 ? 2 : 0) + k;
                float*      p  = a.f + 3 * (size_t)((ci0 + im) * CL + ii);
                atomicAdd(p, fx);
                atomicAdd(p + 1, fy);
                atomicAdd(p + 2, fz);
            }
        }
        if (ENERGY)
        {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                e_el += __shfl_xor_sync(FULL, e_el, d);
                e_lj += __shfl_xor_sync(FULL, e_lj, d);
            }
            if (lane < 2)
            {
                if (a.e_el_f != nullptr)
                {
                    atomicAdd(lane == 0 ? a.e_el_f : a.e_lj_f, lane == 0 ? e_el : e_lj);
                }
                else
                {
                    atomicAdd(a.energies + lane, (double)(lane == 0 ? e_el : e_lj));
                }
            }
        }
    }
}

/* x rvec[natoms] + masked charges -> float4; or the caller's xq with the masked charge put in */
__global__ void fep_nb_pack_kernel(int n, const float* __restrict__ x3, const float4* x4, const float* __restrict__ q,
                                   float4* out) /* x4 may be out */
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        float4 v;
        if (x4 != nullptr)
        {
            v = x4[i];
        }
        else
        {
            v.x = x3[3 * (size_t)i];
            v.y = x3[3 * (size_t)i + 1];
            v.z = x3[3 * (size_t)i + 2];
        }
        v.w    = q[i];
        out[i] = v;
    }
}

/* nbnxn_atomdata_mask_fep (atomdata.cpp:930-964) on the device copies */
__global__ void fep_nb_mask_kernel(int n, const int* __restrict__ atoms, int dummy_type, int* type, float* q)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        type[atoms[i]] = dummy_type;
        q[atoms[i]]    = 0.0f;
    }
}

/* {vc, vvdw} of the handle's accumulator added into the float buffers the fork's nbnxm GPU module reduces from
 * (NBAtomDataGpu::eElec, eLJ) */
__global__ void fep_nb_export_energies_kernel(const double* __restrict__ e, float* eLJ, float* eElec)
{
    if (threadIdx.x == 0 && eElec != nullptr)
    {
        atomicAdd(eElec, (float)e[0]);
    }
    if (threadIdx.x == 1 && eLJ != nullptr)
    {
        atomicAdd(eLJ, (float)e[1]);
    }
}

thread_local std::string g_nb_create_error;

} // namespace

struct fepb200_nb
{
    int            device = 0, sm_count = 148;
    cudaStream_t   own_stream = nullptr, stream = nullptr;
    cudaEvent_t    ev0 = nullptr, ev1 = nullptr;
    std::string    error;
    NbConsts       c{};
    int            vdwmod = 0; /* 0 none / potential shift, 1 force switch, 2 potential switch */
    bool           ewald = false, have_params = false, have_nbfp = false, have_atoms = false, have_list = false;
    int            natoms = 0, ntype = 0, nsci = 0, ncj = 0, nexcl = 0, nitems = 0;
    long long      cluster_pairs = 0, launches = 0;
    int*           d_type = nullptr;
    float*         d_q = nullptr;
    float4*        d_xq = nullptr;
    float*         d_x3 = nullptr;
    float*         d_f = nullptr;
    float2*        d_nbfp = nullptr;
    float*         d_shift = nullptr;
    float*         d_fshift = nullptr;
    double*        d_energies = nullptr;
    fepb200_nb_sci*       d_sci = nullptr;
    fepb200_nb_cj_packed* d_cj = nullptr;
    fepb200_nb_excl*      d_excl = nullptr;
    NbItem*        d_items = nullptr;
    /* fepb200_nb_use_device_list(): the caller's device copies of the list, read instead of ours (NULL: ours) */
    const fepb200_nb_sci*       ext_sci  = nullptr;
    const fepb200_nb_cj_packed* ext_cj   = nullptr;
    const fepb200_nb_excl*      ext_excl = nullptr;
    unsigned int*  d_next = nullptr; /* two heads of the kernel's work queue, used in turn */
    unsigned int   turn = 0;
    float *        e_el_f = nullptr, *e_lj_f = nullptr; /* float energy accumulators of the coming launch (or NULL) */
    size_t         cap_sci = 0, cap_cj = 0, cap_excl = 0, cap_items = 0, cap_atoms = 0;
    float*         h_pinned = nullptr; /* x in / f out staging */
    size_t         cap_pinned = 0;
    float          h_shift[3 * FEPB200_NUM_SHIFT_VECTORS];
    bool           shift_valid = false;
};

namespace
{

int nb_fail(fepb200_nb* h, int code, const std::string& msg)
{
    if (h)
    {
        h->error = msg;
    }
    else
    {
        g_nb_create_error = msg;
    }
    return code;
}

#define NB_CUDA(call)                                                                                  \
    do                                                                                                 \
    {                                                                                                  \
        cudaError_t err_ = (call);                                                                     \
        if (err_ != cudaSuccess)                                                                       \
        {                                                                                              \
            return nb_fail(h, FEPB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err_)); \
        }                                                                                              \
    } while (0)

template<typename T>
int nb_reserve(fepb200_nb* h, T** p, size_t* cap, size_t n)
{
    if (n > *cap)
    {
        if (*p)
        {
            NB_CUDA(cudaFree(*p));
            *p = nullptr;
        }
        const size_t want = n + n / 8 + 64;
        NB_CUDA(cudaMalloc(reinterpret_cast<void**>(p), want * sizeof(T)));
        *cap = want;
    }
    return FEPB200_OK;
}

bool full_electrostatics(int eeltype)
{
    /* usingFullElectrostatics(), md_enums.h:296-314 */
    return (eeltype >= 3 && eeltype <= 6) || (eeltype >= 13 && eeltype <= 15);
}

int nb_launch(fepb200_nb* h, const float4* d_xq, const float* d_shift, int flags, float* d_f, float* d_fshift,
              double* d_energies)
{
    NbArgs a;
    a.xq       = d_xq;
    a.type     = h->d_type;
    a.nbfp     = h->d_nbfp;
    a.shiftvec = d_shift;
    a.sci      = h->ext_sci ? h->ext_sci : h->d_sci;
    a.cj       = reinterpret_cast<const int4*>(h->ext_cj ? h->ext_cj : h->d_cj);
    a.excl     = h->ext_excl ? h->ext_excl : h->d_excl;
    a.items    = h->d_items;
    a.nitems   = h->nitems;
    a.next_item  = h->d_next + (h->turn & 1u);
    a.next_reset = h->d_next + ((h->turn + 1u) & 1u);
    a.e_el_f     = h->e_el_f;
    a.e_lj_f     = h->e_lj_f;
    a.f        = d_f;
    a.fshift   = (flags & FEPB200_DO_SHIFTFORCE) ? d_fshift : nullptr;
    a.energies = d_energies;
    a.c        = h->c;
    const bool energy = (flags & FEPB200_DO_POTENTIAL) != 0;
    NB_CUDA(cudaEventRecord(h->ev0, h->stream));
    if (h->nitems > 0)
    {
        const int per_sm = 4;
        int       grid   = (h->nitems + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        if (grid > h->sm_count * per_sm)
        {
            grid = h->sm_count * per_sm;
        }
        const dim3   block(WARPS_PER_CTA * 32);
        const bool   same   = h->c.rc2 == h->c.rv2;
        const size_t tab    = sizeof(float2) * h->ntype * h->ntype;
        const bool   smem   = tab <= 16384;
        const int    which  = (h->ewald ? 8 : 0) | (energy ? 4 : 0) | (same ? 2 : 0) | (smem ? 1 : 0);
        const size_t shared = WARPS_PER_CTA * ATOMS_SC * (sizeof(float4) + sizeof(int)) + (smem ? tab : 0);
#define NB_CASE(E, V, S, T)                                                                          \
    case ((E) ? 8 : 0) | ((V) ? 4 : 0) | ((S) ? 2 : 0) | ((T) ? 1 : 0):                              \
        if (h->vdwmod == 1)                                                                          \
            fep_nb_kernel<E, V, S, T, 1><<<grid, block, shared, h->stream>>>(a);                     \
        else if (h->vdwmod == 2)                                                                     \
            fep_nb_kernel<E, V, S, T, 2><<<grid, block, shared, h->stream>>>(a);                     \
        else                                                                                         \
            fep_nb_kernel<E, V, S, T, 0><<<grid, block, shared, h->stream>>>(a);                     \
        break;
        switch (which)
        {
            NB_CASE(false, false, false, false)
            NB_CASE(false, false, false, true)
            NB_CASE(false, false, true, false)
            NB_CASE(false, false, true, true)
            NB_CASE(false, true, false, false)
            NB_CASE(false, true, false, true)
            NB_CASE(false, true, true, false)
            NB_CASE(false, true, true, true)
            NB_CASE(true, false, false, false)
            NB_CASE(true, false, false, true)
            NB_CASE(true, false, true, false)
            NB_CASE(true, false, true, true)
            NB_CASE(true, true, false, false)
            NB_CASE(true, true, false, true)
            NB_CASE(true, true, true, false)
            NB_CASE(true, true, true, true)
        }
#undef NB_CASE
        NB_CUDA(cudaGetLastError());
        h->launches++;
        h->turn++;
    }
    NB_CUDA(cudaEventRecord(h->ev1, h->stream));
    return FEPB200_OK;
}

int nb_ready(fepb200_nb* h)
{
    if (!h->have_params || !h->have_nbfp || !h->have_atoms || !h->have_list)
    {
        return nb_fail(h, FEPB200_ERR_STATE, "set_params, set_nbfp, set_atoms and set_pairlist must precede a launch");
    }
    return FEPB200_OK;
}

int nb_upload_shift(fepb200_nb* h, const float* shiftvec)
{
    if (shiftvec == nullptr)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "shiftvec is NULL");
    }
    if (!h->shift_valid || std::memcmp(h->h_shift, shiftvec, sizeof(h->h_shift)) != 0)
    {
        std::memcpy(h->h_shift, shiftvec, sizeof(h->h_shift));
        /* pageable source: the copy is staged by the driver before the call returns */
        NB_CUDA(cudaMemcpyAsync(h->d_shift, h->h_shift, sizeof(h->h_shift), cudaMemcpyHostToDevice, h->stream));
        h->shift_valid = true;
    }
    return FEPB200_OK;
}

} // namespace

extern "C" {

int fepb200_nb_create(fepb200_nb** out, int device_ordinal)
{
    fepb200_nb* h = nullptr;
    if (out == nullptr)
    {
        return nb_fail(nullptr, FEPB200_ERR_INVALID_ARGUMENT, "handle pointer is NULL");
    }
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    {
        return nb_fail(nullptr, FEPB200_ERR_NO_DEVICE, "no CUDA device: libfepb200 has no CPU fallback");
    }
    if (device_ordinal < 0 || device_ordinal >= n)
    {
        return nb_fail(nullptr, FEPB200_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess)
    {
        return nb_fail(nullptr, FEPB200_ERR_CUDA, "cudaGetDeviceProperties failed");
    }
    if (prop.major != 10)
    {
        return nb_fail(nullptr, FEPB200_ERR_NO_DEVICE, "this library is built for sm_100a only");
    }
    h           = new fepb200_nb;
    h->device   = device_ordinal;
    h->sm_count = prop.multiProcessorCount;
    NB_CUDA(cudaSetDevice(device_ordinal));
    NB_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    NB_CUDA(cudaEventCreate(&h->ev0));
    NB_CUDA(cudaEventCreate(&h->ev1));
    NB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_shift), sizeof(h->h_shift)));
    NB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_fshift), sizeof(h->h_shift)));
    NB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_energies), 2 * sizeof(double)));
    NB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_next), 2 * sizeof(unsigned int)));
    NB_CUDA(cudaMemset(h->d_next, 0, 2 * sizeof(unsigned int)));
    *out = h;
    return FEPB200_OK;
}

int fepb200_nb_destroy(fepb200_nb* h)
{
    if (h == nullptr)
    {
        return FEPB200_OK;
    }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_type);
    cudaFree(h->d_q);
    cudaFree(h->d_xq);
    cudaFree(h->d_x3);
    cudaFree(h->d_f);
    cudaFree(h->d_nbfp);
    cudaFree(h->d_shift);
    cudaFree(h->d_fshift);
    cudaFree(h->d_energies);
    cudaFree(h->d_sci);
    cudaFree(h->d_cj);
    cudaFree(h->d_excl);
    cudaFree(h->d_items);
    cudaFree(h->d_next);
    if (h->h_pinned)
    {
        cudaFreeHost(h->h_pinned);
    }
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->own_stream);
    delete h;
    return FEPB200_OK;
}

const char* fepb200_nb_last_error(const fepb200_nb* h)
{
    return h ? h->error.c_str() : g_nb_create_error.c_str();
}

int fepb200_nb_set_stream(fepb200_nb* h, void* stream)
{
    if (h == nullptr)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    NB_CUDA(cudaSetDevice(h->device));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    h->stream = stream ? static_cast<cudaStream_t>(stream) : h->own_stream;
    return FEPB200_OK;
}

int fepb200_nb_set_params(fepb200_nb* h, const fepb200_params* ic)
{
    if (h == nullptr || ic == nullptr)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "NULL argument");
    }
    if (!(ic->rcoulomb > 0.0f) || !(ic->rvdw > 0.0f))
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "cut-offs must be positive");
    }
    h->ewald = full_electrostatics(ic->eeltype);
    NbConsts& c = h->c;
    c.epsfac    = ic->epsfac;
    c.rc2       = ic->rcoulomb * ic->rcoulomb;
    c.rv2       = ic->rvdw * ic->rvdw;
    c.k_rf      = ic->reactionFieldCoefficient;
    c.c_rf      = ic->reactionFieldShift;
    c.sh_ewald  = ic->sh_ewald;
    c.beta      = ic->ewaldcoeff_q;
    c.beta2     = c.beta * c.beta;
    c.beta3     = c.beta2 * c.beta;
    c.disp_cpot = ic->dispersion_shift_cpot;
    c.rep_cpot  = ic->repulsion_shift_cpot;
    /* vdw_modifier: force switch / potential switch select the Lennard-Jones modifiers of the reference's CUDA kernels; every
     * other value gives what nbnxn_kernel_gpu_ref computes (plain LJ, the two cpot constants shift the energy) */
    h->vdwmod     = ic->vdw_modifier == FEPB200_MOD_FORCESWITCH ? 1 : (ic->vdw_modifier == FEPB200_MOD_POTSWITCH ? 2 : 0);
    c.rvdw_switch = ic->rvdw_switch;
    c.dsp_c2 = c.dsp_c3 = c.rep_c2 = c.rep_c3 = c.sw_c3 = c.sw_c4 = c.sw_c5 = 0.0f;
    if (h->vdwmod != 0)
    {
        const double rsw = ic->rvdw_switch, rc = ic->rvdw, d = rc - rsw;
        if (!(rsw >= 0.0) || !(d > 0.0))
        {
            return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "a switched Lennard-Jones needs 0 <= rvdw_switch < rvdw");
        }
        auto fsw = [&](double p, float* c2, float* c3) { /* interaction_const.cpp:216-230 */
            *c2 = (float)(((p + 1) * rsw - (p + 4) * rc) / (std::pow(rc, p + 2) * d * d));
            *c3 = (float)(-((p + 1) * rsw - (p + 3) * rc) / (std::pow(rc, p + 2) * d * d * d));
        };
        fsw(6.0, &c.dsp_c2, &c.dsp_c3);
        fsw(12.0, &c.rep_c2, &c.rep_c3);
        c.sw_c3 = (float)(-10.0 / (d * d * d)); /* interaction_const.cpp:232-245 */
        c.sw_c4 = (float)(15.0 / (d * d * d * d));
        c.sw_c5 = (float)(-6.0 / (d * d * d * d * d));
    }
    if (h->ewald)
    {
        if (c.beta2 * c.rc2 > 16.0f)
        {
            return nb_fail(h, FEPB200_ERR_UNSUPPORTED,
                           "ewaldcoeff_q * rcoulomb > 4 (ewald-rtol below 2e-8): outside the range of the Ewald fits");
        }
        c.self_coef = -c.epsfac * c.beta * 0.5641895835477563f; /* 1/sqrt(pi) */
    }
    else
    {
        c.self_coef = -c.epsfac * 0.5f * c.c_rf;
    }
    h->have_params = true;
    return FEPB200_OK;
}

int fepb200_nb_set_nbfp(fepb200_nb* h, int ntype, const float* nbfp)
{
    if (h == nullptr || nbfp == nullptr || ntype < 1)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "bad nbfp arguments");
    }
    const size_t last = 2 * ((size_t)ntype * ntype - 1);
    for (int t = 0; t < ntype; t++)
    {
        const size_t a = 2 * ((size_t)(ntype - 1) * ntype + t), b = 2 * ((size_t)t * ntype + ntype - 1);
        if (nbfp[a] != 0.0f || nbfp[a + 1] != 0.0f || nbfp[b] != 0.0f || nbfp[b + 1] != 0.0f)
        {
            return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT,
                           "type ntype-1 must be the non-interacting type (all-zero row and column of nbfp)");
        }
    }
    (void)last;
    NB_CUDA(cudaSetDevice(h->device));
    if (h->d_nbfp)
    {
        NB_CUDA(cudaFree(h->d_nbfp));
        h->d_nbfp = nullptr;
    }
    NB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_nbfp), sizeof(float2) * ntype * ntype));
    NB_CUDA(cudaMemcpyAsync(h->d_nbfp, nbfp, sizeof(float2) * ntype * ntype, cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    h->ntype     = ntype;
    h->c.ntype   = ntype;
    h->have_nbfp = true;
    return FEPB200_OK;
}

int fepb200_nb_set_atoms(fepb200_nb* h, int natoms, const int* type, const float* charge)
{
    if (h == nullptr || type == nullptr || charge == nullptr || natoms < 0 || natoms % CL != 0)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "natoms must be a multiple of 8; type and charge must be given");
    }
    if (!h->have_nbfp)
    {
        return nb_fail(h, FEPB200_ERR_STATE, "fepb200_nb_set_nbfp must precede fepb200_nb_set_atoms");
    }
    for (int i = 0; i < natoms; i++)
    {
        if (type[i] < 0 || type[i] >= h->ntype)
        {
            char buf[96];
            snprintf(buf, sizeof(buf), "type[%d] = %d is outside [0, %d)", i, type[i], h->ntype);
            return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, buf);
        }
    }
    NB_CUDA(cudaSetDevice(h->device));
    const size_t n = (size_t)natoms;
    if (n > h->cap_atoms)
    {
        size_t cap = 0;
        for (void** p : { (void**)&h->d_type, (void**)&h->d_q, (void**)&h->d_xq, (void**)&h->d_x3, (void**)&h->d_f })
        {
            if (*p)
            {
                NB_CUDA(cudaFree(*p));
                *p = nullptr;
            }
        }
        cap = n + n / 8 + 64;
        NB_CUDA(cudaMalloc((void**)&h->d_type, cap * sizeof(int)));
        NB_CUDA(cudaMalloc((void**)&h->d_q, cap * sizeof(float)));
        NB_CUDA(cudaMalloc((void**)&h->d_xq, cap * sizeof(float4)));
        NB_CUDA(cudaMalloc((void**)&h->d_x3, cap * 3 * sizeof(float)));
        NB_CUDA(cudaMalloc((void**)&h->d_f, cap * 3 * sizeof(float)));
        h->cap_atoms = cap;
    }
    NB_CUDA(cudaMemcpyAsync(h->d_type, type, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(cudaMemcpyAsync(h->d_q, charge, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    h->natoms     = natoms;
    h->have_atoms = true;
    h->have_list  = false; /* the list indexes these atoms */
    return FEPB200_OK;
}

int fepb200_nb_mask_perturbed(fepb200_nb* h, int n, const int* atoms)
{
    if (h == nullptr || n < 0 || (n > 0 && atoms == nullptr))
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "bad arguments");
    }
    if (!h->have_atoms)
    {
        return nb_fail(h, FEPB200_ERR_STATE, "fepb200_nb_set_atoms must precede fepb200_nb_mask_perturbed");
    }
    for (int i = 0; i < n; i++)
    {
        if (atoms[i] < 0 || atoms[i] >= h->natoms)
        {
            return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "perturbed atom index out of range");
        }
    }
    if (n == 0)
    {
        return FEPB200_OK;
    }
    NB_CUDA(cudaSetDevice(h->device));
    int* d_atoms = nullptr;
    NB_CUDA(cudaMalloc((void**)&d_atoms, n * sizeof(int)));
    NB_CUDA(cudaMemcpyAsync(d_atoms, atoms, n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    fep_nb_mask_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, d_atoms, h->ntype - 1, h->d_type, h->d_q);
    NB_CUDA(cudaGetLastError());
    NB_CUDA(cudaStreamSynchronize(h->stream));
    NB_CUDA(cudaFree(d_atoms));
    h->launches++;
    return FEPB200_OK;
}

int fepb200_nb_get_atoms(const fepb200_nb* hc, int* type, float* charge)
{
    fepb200_nb* h = const_cast<fepb200_nb*>(hc);
    if (h == nullptr || !h->have_atoms)
    {
        return nb_fail(h, FEPB200_ERR_STATE, "no atoms");
    }
    NB_CUDA(cudaSetDevice(h->device));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    if (type)
    {
        NB_CUDA(cudaMemcpy(type, h->d_type, h->natoms * sizeof(int), cudaMemcpyDeviceToHost));
    }
    if (charge)
    {
        NB_CUDA(cudaMemcpy(charge, h->d_q, h->natoms * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return FEPB200_OK;
}

int fepb200_nb_set_pairlist(fepb200_nb* h, int nsci, const fepb200_nb_sci* sci, int ncj, const fepb200_nb_cj_packed* cj,
                            int nexcl, const fepb200_nb_excl* excl)
{
    if (h == nullptr || nsci < 0 || ncj < 0 || nexcl < 1 || (nsci > 0 && sci == nullptr) || (ncj > 0 && cj == nullptr)
        || excl == nullptr)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "bad list arguments (excl[0] must exist)");
    }
    if (!h->have_atoms)
    {
        return nb_fail(h, FEPB200_ERR_STATE, "fepb200_nb_set_atoms must precede fepb200_nb_set_pairlist");
    }
    const int nsuper = h->natoms / (CL * NCL), ncluster = h->natoms / CL;
    if (h->natoms % (CL * NCL) != 0 && nsci > 0)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "natoms must be a multiple of 64 for a super-cluster list");
    }
    /* range checks + the number of listed cluster pairs */
    long long pairs = 0, groups = 0;
    char      buf[160];
    for (int s = 0; s < nsci; s++)
    {
        const fepb200_nb_sci& e = sci[s];
        if (e.sci < 0 || e.sci >= nsuper || e.shift < 0 || e.shift >= FEPB200_NUM_SHIFT_VECTORS || e.cjPackedBegin < 0
            || e.cjPackedEnd < e.cjPackedBegin || e.cjPackedEnd > ncj)
        {
            snprintf(buf, sizeof(buf), "sci[%d] = {%d, %d, %d, %d} is out of range", s, e.sci, e.shift, e.cjPackedBegin,
                     e.cjPackedEnd);
            return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, buf);
        }
        groups += e.cjPackedEnd - e.cjPackedBegin;
    }
    for (int g = 0; g < ncj; g++)
    {
        const fepb200_nb_cj_packed& p = cj[g];
        for (int jm = 0; jm < FEPB200_NB_JGROUP_SIZE; jm++)
        {
            if (p.cj[jm] < 0 || p.cj[jm] >= ncluster)
            {
                snprintf(buf, sizeof(buf), "cj[%d].cj[%d] = %d is outside [0, %d)", g, jm, p.cj[jm], ncluster);
                return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, buf);
            }
        }
        for (int half = 0; half < FEPB200_NB_CLUSTERPAIR_SPLIT; half++)
        {
            if (p.imei[half].excl_ind < 0 || p.imei[half].excl_ind >= nexcl)
            {
                snprintf(buf, sizeof(buf), "cj[%d].imei[%d].excl_ind = %d is outside [0, %d)", g, half,
                         p.imei[half].excl_ind, nexcl);
                return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, buf);
            }
        }
        pairs += __builtin_popcount(p.imei[0].imask);
    }
    /* work items: an entry's packed j-cluster entries in chunks, sized so that every resident warp gets several */
    const long long resident = (long long)h->sm_count * 4 * WARPS_PER_CTA;
    long long       chunk    = groups / (resident * 4);
    chunk                    = chunk < 2 ? 2 : (chunk > 16 ? 16 : chunk);
    std::vector<NbItem> items;
    items.reserve((size_t)(groups / chunk + nsci));
    for (int s = 0; s < nsci; s++)
    {
        const fepb200_nb_sci& e = sci[s];
        /* the entry whose first j-cluster is the super-cluster's own first cluster, central cell, books the self term */
        const bool self = e.cjPackedEnd > e.cjPackedBegin && e.shift == CENTRAL_SHIFT
                          && cj[e.cjPackedBegin].cj[0] == e.sci * NCL;
        if (e.cjPackedEnd == e.cjPackedBegin)
        {
            continue;
        }
        for (int b = e.cjPackedBegin; b < e.cjPackedEnd; b += (int)chunk)
        {
            const int end = b + (int)chunk < e.cjPackedEnd ? b + (int)chunk : e.cjPackedEnd;
            items.push_back(NbItem{ s, b, end, (self && b == e.cjPackedBegin) ? 1 : 0 });
        }
    }
    NB_CUDA(cudaSetDevice(h->device));
    int rc;
    if ((rc = nb_reserve(h, &h->d_sci, &h->cap_sci, (size_t)nsci)) != FEPB200_OK
        || (rc = nb_reserve(h, &h->d_cj, &h->cap_cj, (size_t)ncj)) != FEPB200_OK
        || (rc = nb_reserve(h, &h->d_excl, &h->cap_excl, (size_t)nexcl)) != FEPB200_OK
        || (rc = nb_reserve(h, &h->d_items, &h->cap_items, items.size())) != FEPB200_OK)
    {
        return rc;
    }
    NB_CUDA(cudaMemcpyAsync(h->d_sci, sci, sizeof(fepb200_nb_sci) * nsci, cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(cudaMemcpyAsync(h->d_cj, cj, sizeof(fepb200_nb_cj_packed) * ncj, cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(cudaMemcpyAsync(h->d_excl, excl, sizeof(fepb200_nb_excl) * nexcl, cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(cudaMemcpyAsync(h->d_items, items.data(), sizeof(NbItem) * items.size(), cudaMemcpyHostToDevice, h->stream));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    h->nsci          = nsci;
    h->ncj           = ncj;
    h->nexcl         = nexcl;
    h->nitems        = (int)items.size();
    h->cluster_pairs = pairs;
    h->have_list     = true;
    h->ext_sci       = nullptr;
    h->ext_cj        = nullptr;
    h->ext_excl      = nullptr;
    return FEPB200_OK;
}

int fepb200_nb_launch_device(fepb200_nb* h, const float* d_xq, const float* shiftvec, int flags, float* d_f,
                             float* d_fshift, double* d_energies)
{
    if (h == nullptr || d_xq == nullptr || d_f == nullptr)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "d_xq and d_f must be given");
    }
    if ((flags & FEPB200_DO_SHIFTFORCE) && d_fshift == nullptr)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "an output selected by flags is NULL");
    }
    int rc = nb_ready(h);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    NB_CUDA(cudaSetDevice(h->device));
    const float* d_shift = h->d_shift;
    if (flags & FEPB200_NB_SHIFTVEC_ON_DEVICE)
    {
        if (shiftvec == nullptr)
        {
            return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "shiftvec is NULL");
        }
        d_shift = shiftvec;
    }
    else if ((rc = nb_upload_shift(h, shiftvec)) != FEPB200_OK)
    {
        return rc;
    }
    if ((flags & FEPB200_DO_POTENTIAL) && d_energies == nullptr)
    {
        /* the handle's own accumulator, for fepb200_nb_export_energies_device() */
        NB_CUDA(cudaMemsetAsync(h->d_energies, 0, 2 * sizeof(double), h->stream));
        d_energies = h->d_energies;
    }
    const float4* xq = reinterpret_cast<const float4*>(d_xq);
    if (!(flags & FEPB200_NB_Q_FROM_XQ))
    {
        fep_nb_pack_kernel<<<(h->natoms + 255) / 256, 256, 0, h->stream>>>(h->natoms, nullptr, xq, h->d_q, h->d_xq);
        NB_CUDA(cudaGetLastError());
        h->launches++;
        xq = h->d_xq;
    }
    return nb_launch(h, xq, d_shift, flags, d_f, d_fshift, d_energies);
}

int fepb200_nb_launch_device_float_energies(fepb200_nb* h, const float* d_xq, const float* shiftvec, int flags, float* d_f,
                                            float* d_fshift, float* d_eLJ, float* d_eElec)
{
    if (h == nullptr)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    if ((flags & FEPB200_DO_POTENTIAL) && (d_eLJ == nullptr || d_eElec == nullptr))
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "an output selected by flags is NULL");
    }
    h->e_el_f    = (flags & FEPB200_DO_POTENTIAL) ? d_eElec : nullptr;
    h->e_lj_f    = (flags & FEPB200_DO_POTENTIAL) ? d_eLJ : nullptr;
    /* the double accumulator argument is unused on this route: any non-NULL pointer keeps fepb200_nb_launch_device from
     * clearing and selecting the handle's own */
    const int rc = fepb200_nb_launch_device(h, d_xq, shiftvec, flags, d_f, d_fshift, h->d_energies);
    h->e_el_f    = nullptr;
    h->e_lj_f    = nullptr;
    return rc;
}

int fepb200_nb_use_device_list(fepb200_nb* h, const fepb200_nb_sci* d_sci, const fepb200_nb_cj_packed* d_cj,
                               const fepb200_nb_excl* d_excl)
{
    if (h == nullptr || !h->have_list)
    {
        return nb_fail(h, FEPB200_ERR_STATE, "fepb200_nb_set_pairlist must precede fepb200_nb_use_device_list");
    }
    if ((d_sci == nullptr) != (d_cj == nullptr) || (d_sci == nullptr) != (d_excl == nullptr))
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "give all three device arrays, or none");
    }
    h->ext_sci  = d_sci;
    h->ext_cj   = d_cj;
    h->ext_excl = d_excl;
    return FEPB200_OK;
}

int fepb200_nb_export_energies_device(fepb200_nb* h, float* d_eLJ, float* d_eElec)
{
    if (h == nullptr)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    NB_CUDA(cudaSetDevice(h->device));
    fep_nb_export_energies_kernel<<<1, 32, 0, h->stream>>>(h->d_energies, d_eLJ, d_eElec);
    NB_CUDA(cudaGetLastError());
    h->launches++;
    return FEPB200_OK;
}

static int nb_compute_host(fepb200_nb* h, const float* x, int xstride, const float* shiftvec, int flags, float* f,
                           float* fshift, double* vc, double* vvdw)
{
    if (h == nullptr || x == nullptr || f == nullptr)
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "x and f must be given");
    }
    if (((flags & FEPB200_DO_SHIFTFORCE) && fshift == nullptr) || ((flags & FEPB200_DO_POTENTIAL) && (vc == nullptr || vvdw == nullptr)))
    {
        return nb_fail(h, FEPB200_ERR_INVALID_ARGUMENT, "an output selected by flags is NULL");
    }
    int rc = nb_ready(h);
    if (rc != FEPB200_OK)
    {
        return rc;
    }
    NB_CUDA(cudaSetDevice(h->device));
    if ((rc = nb_upload_shift(h, shiftvec)) != FEPB200_OK)
    {
        return rc;
    }
    const size_t n3    = 3 * (size_t)h->natoms;
    const size_t nx    = (size_t)xstride * h->natoms;
    const size_t words = nx + n3 + 3 * FEPB200_NUM_SHIFT_VECTORS + 4;
    if (words > h->cap_pinned)
    {
        if (h->h_pinned)
        {
            NB_CUDA(cudaFreeHost(h->h_pinned));
            h->h_pinned = nullptr;
        }
        NB_CUDA(cudaMallocHost((void**)&h->h_pinned, words * sizeof(float)));
        h->cap_pinned = words;
    }
    /* pinned staging block: {vc, vvdw} as doubles first (alignment), then x in, f out, shift forces out */
    double* he  = reinterpret_cast<double*>(h->h_pinned);
    float*  hx  = h->h_pinned + 4;
    float*  hf  = hx + nx;
    float*  hfs = hf + n3;
    std::memcpy(hx, x, nx * sizeof(float));
    if (xstride == 3)
    {
        NB_CUDA(cudaMemcpyAsync(h->d_x3, hx, nx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        fep_nb_pack_kernel<<<(h->natoms + 255) / 256, 256, 0, h->stream>>>(h->natoms, h->d_x3, nullptr, h->d_q, h->d_xq);
        NB_CUDA(cudaGetLastError());
        h->launches++;
    }
    else
    {
        NB_CUDA(cudaMemcpyAsync(h->d_xq, hx, nx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        if (!(flags & FEPB200_NB_Q_FROM_XQ))
        {
            fep_nb_pack_kernel<<<(h->natoms + 255) / 256, 256, 0, h->stream>>>(h->natoms, nullptr, h->d_xq, h->d_q, h->d_xq);
            NB_CUDA(cudaGetLastError());
            h->launches++;
        }
    }
    NB_CUDA(cudaMemsetAsync(h->d_f, 0, n3 * sizeof(float), h->stream));
    NB_CUDA(cudaMemsetAsync(h->d_fshift, 0, sizeof(h->h_shift), h->stream));
    NB_CUDA(cudaMemsetAsync(h->d_energies, 0, 2 * sizeof(double), h->stream));
    if ((rc = nb_launch(h, h->d_xq, h->d_shift, flags, h->d_f, h->d_fshift, (flags & FEPB200_DO_POTENTIAL) ? h->d_energies : nullptr))
        != FEPB200_OK)
    {
        return rc;
    }
    NB_CUDA(cudaMemcpyAsync(hf, h->d_f, n3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (flags & FEPB200_DO_SHIFTFORCE)
    {
        NB_CUDA(cudaMemcpyAsync(hfs, h->d_fshift, sizeof(h->h_shift), cudaMemcpyDeviceToHost, h->stream));
    }
    if (flags & FEPB200_DO_POTENTIAL)
    {
        NB_CUDA(cudaMemcpyAsync(he, h->d_energies, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    NB_CUDA(cudaStreamSynchronize(h->stream));
    const bool clear = (flags & FEPB200_CLEAR_OUTPUTS) != 0;
    if (clear)
    {
        std::memcpy(f, hf, n3 * sizeof(float));
    }
    else
    {
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < (long long)n3; k++)
        {
            f[k] += hf[k];
        }
    }
    if (flags & FEPB200_DO_SHIFTFORCE)
    {
        for (int k = 0; k < 3 * FEPB200_NUM_SHIFT_VECTORS; k++)
        {
            fshift[k] = clear ? hfs[k] : fshift[k] + hfs[k];
        }
    }
    if (flags & FEPB200_DO_POTENTIAL)
    {
        *vc   = clear ? he[0] : *vc + he[0];
        *vvdw = clear ? he[1] : *vvdw + he[1];
    }
    return FEPB200_OK;
}

int fepb200_nb_compute(fepb200_nb* h, const float* x, const float* shiftvec, int flags, float* f, float* fshift, double* vc,
                       double* vvdw)
{
    return nb_compute_host(h, x, 3, shiftvec, flags & ~FEPB200_NB_Q_FROM_XQ, f, fshift, vc, vvdw);
}

int fepb200_nb_compute_xyzq(fepb200_nb* h, const float* xq, const float* shiftvec, int flags, float* f, float* fshift,
                            double* vc, double* vvdw)
{
    return nb_compute_host(h, xq, 4, shiftvec, flags, f, fshift, vc, vvdw);
}

int fepb200_nb_wait(fepb200_nb* h)
{
    if (h == nullptr)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    NB_CUDA(cudaSetDevice(h->device));
    NB_CUDA(cudaStreamSynchronize(h->stream));
    return FEPB200_OK;
}

long long fepb200_nb_launch_count(const fepb200_nb* h)
{
    return h ? h->launches : 0;
}

int fepb200_nb_last_kernel_ms(fepb200_nb* h, float* ms)
{
    if (h == nullptr || ms == nullptr)
    {
        return FEPB200_ERR_INVALID_ARGUMENT;
    }
    NB_CUDA(cudaSetDevice(h->device));
    NB_CUDA(cudaEventSynchronize(h->ev1));
    NB_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return FEPB200_OK;
}

long long fepb200_nb_cluster_pairs(const fepb200_nb* h)
{
    return h ? h->cluster_pairs : 0;
}

} /* extern "C" */
