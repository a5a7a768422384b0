/*
 * tests/shim_standin/fepb200_nb_standin.c -- TEST INFRASTRUCTURE: a CPU stand-in for the fepb200_nb_* entry points that
 * integration/gromacs_shim/fepb200_nb_shim.h binds, answered by the CPU checker (oracle/nb_oracle.c), so that the SHIM --
 * what it hands over, when, and where it adds the results -- can be tested inside the reference's mdrun without a GPU
 * (tests/test_nb_shim_cpu.py).  Linked into libfepb200_standin.so next to fepb200_standin.c.  Not a product path:
 * libfepb200.so has no CPU route.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fepb200_nb.h"

/* oracle/nb_oracle.c */
typedef struct
{
    int    eeltype;
    double epsfac, rcoulomb, rvdw, rlist, k_rf, c_rf, sh_ewald, beta, disp_cpot, rep_cpot, min_rsq, tab_scale;
    int    tab_size;
    const double* tableF;
    int    vdw_switch_kind;
    double rvdw_switch;
} nbo_params;
int nbo_run(int natoms, const double* xq, const int* type, int ntype, const double* nbfp, const nbo_params* p, int nsci,
            const void* sci, int ncj, const void* cj, int nexcl, const void* excl, const double* shiftvec, int want_energy,
            double* f, double* fshift, double* vc, double* vvdw);

struct fepb200_nb
{
    nbo_params p;
    int        ntype, natoms, nsci, ncj, nexcl;
    double*    nbfp;
    int*       type;
    float*     q;
    void *     sci, *cj, *excl;
    long       n_compute, n_list, n_atoms, n_params;
    char       err[256];
};

static int trace(void)
{
    return getenv("FEPB200_STANDIN_TRACE") != NULL;
}

int fepb200_nb_create(fepb200_nb** h, int device_ordinal)
{
    (void)device_ordinal;
    *h = (fepb200_nb*)calloc(1, sizeof(fepb200_nb));
    fprintf(stderr, "fepb200_nb CPU STAND-IN (tests only)\n");
    return FEPB200_OK;
}

int fepb200_nb_destroy(fepb200_nb* h)
{
    free(h);
    return FEPB200_OK;
}

const char* fepb200_nb_last_error(const fepb200_nb* h)
{
    return h ? h->err : "";
}

int fepb200_nb_set_params(fepb200_nb* h, const fepb200_params* ic)
{
    memset(&h->p, 0, sizeof(h->p));
    h->p.eeltype   = ic->eeltype;
    h->p.epsfac    = ic->epsfac;
    h->p.rcoulomb  = ic->rcoulomb;
    h->p.rvdw      = ic->rvdw;
    h->p.k_rf      = ic->reactionFieldCoefficient;
    h->p.c_rf      = ic->reactionFieldShift;
    h->p.sh_ewald  = ic->sh_ewald;
    h->p.beta      = ic->ewaldcoeff_q;
    h->p.disp_cpot = ic->dispersion_shift_cpot;
    h->p.rep_cpot  = ic->repulsion_shift_cpot;
    h->p.min_rsq   = FEPB200_NB_MIN_RSQ;
    h->n_params++;
    return FEPB200_OK;
}

int fepb200_nb_set_nbfp(fepb200_nb* h, int ntype, const float* nbfp)
{
    free(h->nbfp);
    h->ntype = ntype;
    h->nbfp  = (double*)malloc(sizeof(double) * 2 * ntype * ntype);
    for (int k = 0; k < 2 * ntype * ntype; k++)
    {
        h->nbfp[k] = nbfp[k];
    }
    return FEPB200_OK;
}

int fepb200_nb_set_atoms(fepb200_nb* h, int natoms, const int* type, const float* charge)
{
    free(h->type);
    free(h->q);
    h->natoms = natoms;
    h->type   = (int*)malloc(sizeof(int) * (natoms + 1));
    h->q      = (float*)malloc(sizeof(float) * (natoms + 1));
    memcpy(h->type, type, sizeof(int) * natoms);
    memcpy(h->q, charge, sizeof(float) * natoms);
    h->n_atoms++;
    return FEPB200_OK;
}

int fepb200_nb_set_pairlist(fepb200_nb* h, int nsci, const fepb200_nb_sci* sci, int ncj, const fepb200_nb_cj_packed* cj,
                            int nexcl, const fepb200_nb_excl* excl)
{
    free(h->sci);
    free(h->cj);
    free(h->excl);
    h->nsci = nsci;
    h->ncj  = ncj;
    h->nexcl = nexcl;
    h->sci  = malloc(sizeof(*sci) * (nsci + 1));
    h->cj   = malloc(sizeof(*cj) * (ncj + 1));
    h->excl = malloc(sizeof(*excl) * (nexcl + 1));
    memcpy(h->sci, sci, sizeof(*sci) * nsci);
    memcpy(h->cj, cj, sizeof(*cj) * ncj);
    memcpy(h->excl, excl, sizeof(*excl) * nexcl);
    h->n_list++;
    return FEPB200_OK;
}

int fepb200_nb_compute_xyzq(fepb200_nb* h, const float* xq, const float* shiftvec, int flags, float* f, float* fshift,
                            double* vc, double* vvdw)
{
    const int n   = h->natoms;
    double*   x   = (double*)malloc(sizeof(double) * 4 * (n + 1));
    double*   fo  = (double*)malloc(sizeof(double) * 3 * (n + 1));
    double    sv[135], fs[135], evc = 0, evv = 0;
    for (int i = 0; i < n; i++)
    {
        x[4 * i]     = xq[4 * i];
        x[4 * i + 1] = xq[4 * i + 1];
        x[4 * i + 2] = xq[4 * i + 2];
        x[4 * i + 3] = (flags & FEPB200_NB_Q_FROM_XQ) ? xq[4 * i + 3] : h->q[i];
    }
    for (int k = 0; k < 135; k++)
    {
        sv[k] = shiftvec[k];
    }
    nbo_run(n, x, h->type, h->ntype, h->nbfp, &h->p, h->nsci, h->sci, h->ncj, h->cj, h->nexcl, h->excl, sv,
            (flags & FEPB200_DO_POTENTIAL) != 0, fo, fs, &evc, &evv);
    const int clear = (flags & FEPB200_CLEAR_OUTPUTS) != 0;
    for (int k = 0; k < 3 * n; k++)
    {
        f[k] = (float)((clear ? 0.0 : (double)f[k]) + fo[k]);
    }
    if (flags & FEPB200_DO_SHIFTFORCE)
    {
        for (int k = 0; k < 135; k++)
        {
            fshift[k] = (float)((clear ? 0.0 : (double)fshift[k]) + fs[k]);
        }
    }
    if (flags & FEPB200_DO_POTENTIAL)
    {
        *vc   = (clear ? 0.0 : *vc) + evc;
        *vvdw = (clear ? 0.0 : *vvdw) + evv;
    }
    free(x);
    free(fo);
    h->n_compute++;
    if (trace())
    {
        fprintf(stderr, "nb standin: compute %ld set_pairlist %ld set_atoms %ld set_params %ld flags 0x%x\n", h->n_compute,
                h->n_list, h->n_atoms, h->n_params, flags);
    }
    return FEPB200_OK;
}
