/*
 * oracle/nb_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C, double-precision restatement of the reference's kernel for GPU-layout cluster pair lists --
 * the non-perturbed neighbour of the FEP path:
 *   nbnxn_kernel_gpu_ref()        src/gromacs/nbnxm/kernels_reference/kernel_gpu_ref.cpp:54-354
 *   nbnxn_atomdata_mask_fep()     src/gromacs/nbnxm/atomdata.cpp:930-964
 * List structures: nbnxn_sci_t, nbnxn_cj_packed_t, nbnxn_excl_t (src/gromacs/nbnxm/pairlist.h:195-280),
 * 8-atom clusters, 8 clusters per super-cluster, 4 j-clusters per packed entry, 2 half-warps per cluster pair
 * (pairlistparams.h:63-98).
 *
 * PINNED: tests/test_oracle_nb.py compares it with oracle/_ref/libnbref_dp.so, which is that very reference file
 * compiled in place (oracle/ref_build/nb_harness.cpp), on random systems, with the same Ewald force table.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.  The product never does.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CL 8      /* atoms per cluster          pairlistparams.h:65  */
#define NCL_SC 8  /* clusters per super-cluster pairlist.h:174        */
#define JGROUP 4  /* j-clusters per packed entry pairlist.h:180       */
#define NSHIFT 45
#define CENTRAL_SHIFT 22

typedef struct
{
    int sci, shift, cj_begin, cj_end;
} nbo_sci; /* nbnxn_sci_t */
typedef struct
{
    int cj[JGROUP];
    struct
    {
        unsigned imask;
        int      excl_ind;
    } half[2];
} nbo_cj; /* nbnxn_cj_packed_t */
typedef struct
{
    unsigned pair[32];
} nbo_excl; /* nbnxn_excl_t */

typedef struct
{
    int    eeltype; /* 0 cut, 1 RF, 3 PME, 4 Ewald ... : "full electrostatics" = PME/Ewald family */
    double epsfac, rcoulomb, rvdw, rlist;
    double k_rf, c_rf;
    double sh_ewald, beta;
    double disp_cpot, rep_cpot;
    double min_rsq;   /* c_nbnxnMinDistanceSquared of the build being mirrored (pairlist.h:162,167) */
    double tab_scale; /* Ewald force table, linear interpolation (kernel_gpu_ref.cpp:239-246); */
    int    tab_size;  /* tab_size == 0: the analytical function the table samples                */
    const double* tableF;
    /* Lennard-Jones modifiers of the reference's CUDA kernels, which nbnxn_kernel_gpu_ref does not have
     * (nbnxm/cuda/nbnxm_cuda_kernel_utils.cuh:104-211, applied at nbnxm_cuda_kernel.cuh:560-590):
     * 0 none / potential shift only, 1 force switch, 2 potential switch */
    int    vdw_switch_kind;
    double rvdw_switch;
} nbo_params;

/* force-switch constants for r^-p (mdtypes/interaction_const.cpp:216-230) */
void nbo_force_switch_constants(double p, double rsw, double rc, double* c2, double* c3, double* cpot)
{
    const double d = rc - rsw;
    *c2            = ((p + 1) * rsw - (p + 4) * rc) / (pow(rc, p + 2) * d * d);
    *c3            = -((p + 1) * rsw - (p + 3) * rc) / (pow(rc, p + 2) * d * d * d);
    *cpot          = -pow(rc, -p) + p * *c2 / 3 * d * d * d + p * *c3 / 4 * d * d * d * d;
}

static int full_electrostatics(int eeltype)
{
    /* usingFullElectrostatics() (md_enums.h:296-314): Pme 3, Ewald 4, P3mAD 5, Poisson 6, PmeUser 13, PmeSwitch 14,
     * PmeUserSwitch 15 */
    return (eeltype >= 3 && eeltype <= 6) || (eeltype >= 13 && eeltype <= 15);
}

/* -d/dr [erf(beta r)/r]: what the reference's table holds at r = i/scale */
double nbo_ewald_force_lr(double beta, double r)
{
    if (r < 1e-9)
    {
        return 0.0;
    }
    const double br = beta * r;
    return erf(br) / (r * r) - 2.0 * beta / sqrt(M_PI) * exp(-br * br) / r;
}

void nbo_fill_table(double beta, double scale, int n, double* tab)
{
    for (int i = 0; i < n; i++)
    {
        tab[i] = nbo_ewald_force_lr(beta, i / scale);
    }
}

/* atomdata.cpp:930-964: a perturbed atom keeps its place in the cluster grid but stops interacting in the
 * cluster kernel: type -> the last (all-zero) type, charge -> 0.  Its interactions are the FEP kernel's. */
void nbo_mask_perturbed(double* xq, int* type, int ntype, int n, const int* atoms)
{
    for (int k = 0; k < n; k++)
    {
        type[atoms[k]]       = ntype - 1;
        xq[4 * atoms[k] + 3] = 0.0;
    }
}

int nbo_run(int natoms, const double* xq, const int* type, int ntype, const double* nbfp, const nbo_params* p,
            int nsci, const nbo_sci* sci, int ncj, const nbo_cj* cj, int nexcl, const nbo_excl* excl,
            const double* shiftvec, int want_energy, double* f, double* fshift, double* vc, double* vvdw)
{
    (void)ncj;
    (void)nexcl;
    const int    ewald = full_electrostatics(p->eeltype);
    const double rc2 = p->rcoulomb * p->rcoulomb, rv2 = p->rvdw * p->rvdw;
    memset(f, 0, sizeof(double) * 3 * (size_t)natoms);
    memset(fshift, 0, sizeof(double) * 3 * NSHIFT);
    *vc = *vvdw = 0.0;

    for (int s = 0; s < nsci; s++)
    {
        const nbo_sci* e  = &sci[s];
        const double*  sh = shiftvec + 3 * e->shift;
        double         e_el = 0.0, e_lj = 0.0;

        /* self term of the charges of a super-cluster, booked with the entry that starts at its own first
         * cluster in the central cell (kernel_gpu_ref.cpp:122-147) */
        if (e->shift == CENTRAL_SHIFT && cj[e->cj_begin].cj[0] == e->sci * NCL_SC)
        {
            double q2 = 0.0;
            for (int a = e->sci * NCL_SC * CL; a < (e->sci + 1) * NCL_SC * CL; a++)
            {
                q2 += xq[4 * a + 3] * xq[4 * a + 3];
            }
            e_el = ewald ? -p->epsfac * p->beta / sqrt(M_PI) * q2 : -p->epsfac * 0.5 * p->c_rf * q2;
        }

        for (int g = e->cj_begin; g < e->cj_end; g++)
        {
            const nbo_cj* grp = &cj[g];
            for (int slot = 0; slot < JGROUP * NCL_SC; slot++) /* bit = jm * 8 + im */
            {
                if (!((grp->half[0].imask >> slot) & 1u))
                {
                    continue;
                }
                const int cjn = grp->cj[slot / NCL_SC];
                const int cin = e->sci * NCL_SC + slot % NCL_SC;
                for (int ii = 0; ii < CL; ii++)
                {
                    const int    ia = cin * CL + ii;
                    const double xi = xq[4 * ia] + sh[0], yi = xq[4 * ia + 1] + sh[1], zi = xq[4 * ia + 2] + sh[2];
                    const double qi = p->epsfac * xq[4 * ia + 3];
                    double       fi[3] = { 0, 0, 0 };
                    for (int jj = 0; jj < CL; jj++)
                    {
                        const int ja = cjn * CL + jj;
                        /* a cluster against itself in the central cell: upper triangle only */
                        if (e->shift == CENTRAL_SHIFT && cin == cjn && ja <= ia)
                        {
                            continue;
                        }
                        /* interaction bit: half jj/4 of the cluster pair, word (jj%4)*8 + ii, bit `slot` */
                        const nbo_excl* x   = &excl[grp->half[jj / (CL / 2)].excl_ind];
                        const double    bit = (double)((x->pair[(jj % (CL / 2)) * CL + ii] >> slot) & 1u);

                        const double dx = xi - xq[4 * ja], dy = yi - xq[4 * ja + 1], dz = zi - xq[4 * ja + 2];
                        double       r2 = dx * dx + dy * dy + dz * dz;
                        if (r2 >= rc2)
                        {
                            continue;
                        }
                        if (r2 < p->min_rsq)
                        {
                            r2 = p->min_rsq;
                        }
                        const double rinv = 1.0 / sqrt(r2), rinv2 = rinv * rinv;
                        const double qq = qi * xq[4 * ja + 3];
                        double       fs, v_el;
                        if (!ewald)
                        {
                            const double kr2 = p->k_rf * r2;
                            fs               = qq * (bit * rinv - 2.0 * kr2) * rinv2;
                            v_el             = qq * (bit * rinv + kr2 - p->c_rf);
                        }
                        else
                        {
                            const double r = r2 * rinv;
                            double       flr;
                            if (p->tab_size > 0)
                            {
                                const double t  = r * p->tab_scale;
                                const int    n0 = (int)t;
                                const double w  = t - n0;
                                flr             = (1.0 - w) * p->tableF[n0] + w * p->tableF[n0 + 1];
                            }
                            else
                            {
                                flr = nbo_ewald_force_lr(p->beta, r);
                            }
                            fs   = qq * (bit * rinv2 - flr) * rinv;
                            v_el = qq * ((bit - erf(p->beta * r)) * rinv - bit * p->sh_ewald);
                        }
                        if (r2 < rv2)
                        {
                            const double* c   = nbfp + 2 * (ntype * type[ia] + type[ja]);
                            const double  r6  = bit * rinv2 * rinv2 * rinv2;
                            const double  v6  = c[0] * r6;
                            const double  v12 = c[1] * r6 * r6;
                            double        f_lj = (v12 - v6) * rinv2; /* force / r */
                            double        v_lj = (v12 + bit * c[1] * p->rep_cpot) / 12.0 - (v6 + bit * c[0] * p->disp_cpot) / 6.0;
                            if (p->vdw_switch_kind != 0)
                            {
                                /* c[0] = 6 C6, c[1] = 12 C12 are what the CUDA kernel calls c6, c12; like there, the
                                 * switch terms are not multiplied by the interaction bit */
                                const double r = r2 * rinv;
                                double       sd = r - p->rvdw_switch;
                                sd              = sd > 0.0 ? sd : 0.0;
                                if (p->vdw_switch_kind == 1)
                                {
                                    double d2, d3, q2, q3, cp;
                                    nbo_force_switch_constants(6.0, p->rvdw_switch, p->rvdw, &d2, &d3, &cp);
                                    nbo_force_switch_constants(12.0, p->rvdw_switch, p->rvdw, &q2, &q3, &cp);
                                    f_lj += (-c[0] * (d2 + d3 * sd) + c[1] * (q2 + q3 * sd)) * sd * sd * rinv;
                                    v_lj += (c[0] * (d2 / 3 + d3 / 4 * sd) - c[1] * (q2 / 3 + q3 / 4 * sd)) * sd * sd * sd;
                                }
                                else
                                {
                                    const double d   = p->rvdw - p->rvdw_switch;
                                    const double c3  = -10.0 / (d * d * d), c4 = 15.0 / (d * d * d * d), c5 = -6.0 / (d * d * d * d * d);
                                    const double sw  = 1.0 + (c3 + (c4 + c5 * sd) * sd) * sd * sd * sd;
                                    const double dsw = (3 * c3 + (4 * c4 + 5 * c5 * sd) * sd) * sd * sd;
                                    f_lj             = f_lj * sw - rinv * v_lj * dsw;
                                    v_lj *= sw;
                                }
                            }
                            fs += f_lj;
                            /* the reference books the Coulomb energy only for pairs inside the LJ cut-off
                             * (kernel_gpu_ref.cpp:264-287: `vctot += vcoul` sits in the `rsq < rvdw2` branch) */
                            e_el += v_el;
                            e_lj += v_lj;
                        }
                        fi[0] += fs * dx;
                        fi[1] += fs * dy;
                        fi[2] += fs * dz;
                        f[3 * ja] -= fs * dx;
                        f[3 * ja + 1] -= fs * dy;
                        f[3 * ja + 2] -= fs * dz;
                    }
                    for (int d = 0; d < 3; d++)
                    {
                        f[3 * ia + d] += fi[d];
                        fshift[3 * e->shift + d] += fi[d];
                    }
                }
            }
        }
        if (want_energy)
        {
            *vc += e_el;
            *vvdw += e_lj;
        }
    }
    return 0;
}
