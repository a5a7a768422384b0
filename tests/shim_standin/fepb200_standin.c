/*
 * tests/shim_standin/fepb200_standin.c -- TEST INFRASTRUCTURE, not a product path and not a CPU
 * fallback: a stand-in for libfepb200.so that exports the nine entry points the reference-side
 * shim binds (integration/gromacs_shim/fepb200_shim.h) and answers them with the CPU oracle
 * (oracle/fep_oracle.c).  Its only use is tests/test_shim_cpu.py: running the reference's patched
 * mdrun in a container without a GPU, to check what the SHIM does -- which arrays it hands over and
 * when (search-step cadence), flag assembly, result routing -- against the reference's own CPU
 * route.  The real library has no such path: without a CUDA device fepb200_create() fails.
 *
 * Semantics follow include/fepb200.h: inputs are copied at set_* time, compute() accumulates (+=).
 *
 * For tests/test_gpu_shim_cpu.py it also answers the five device-resident entry points the GPU-route shim
 * binds (fepb200_set_stream, fepb200_gather_xq_device, fepb200_launch, fepb200_add_forces_device,
 * fepb200_export_scalars_device) on host arrays.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fepb200.h"

/* oracle/fep_oracle.c */
typedef struct fep_oracle_params
{
    int    eeltype, vdwtype, vdw_modifier;
    double epsfac, rcoulomb, rvdw, rvdw_switch, krf, crf;
    double sh_ewald, sh_lj_ewald, ewaldcoeff_q, ewaldcoeff_lj, dispersion_cpot, repulsion_cpot;
    int    softcoreType;
    double alphaVdw, alphaCoulomb;
    int    lambdaPower;
    double sigma6WithInvalidSigma, sigma6Minimum, gapsysScaleVdW, gapsysScaleCoul, gapsysSigma6VdW;
} fep_oracle_params;
int fep_oracle_dispatch(const fep_oracle_params* p, int use_simd, int nthreads, int ntype, const double* nbfp,
                        const double* nbfp_grid, int natoms, const double* x, const double* qA, const double* qB,
                        const int* typeA, const int* typeB, const double* shiftvec, int nri, const int* iinr,
                        const int* gid, const int* shift, const int* jindex, const int* jjnr, const int* excl, int ngrp,
                        int flags, const double* lambda, int nforeign, const double* all_lambda_coul,
                        const double* all_lambda_vdw, double* f, double* fshift, double* Vc, double* Vv, double* dvdl,
                        double* foreign_e, double* foreign_dvdl, int repeats, double* seconds);

struct fepb200_ctx
{
    fep_oracle_params p;
    int               have_params, ntype, natoms, nri, nrj, ngrp, nforeign;
    double *          nbfp, *nbfp_grid, *qA, *qB, *all_c, *all_v;
    int *             typeA, *typeB, *iinr, *gid, *shift, *jindex, *jjnr, *excl;
    double            lambda[FEPB200_NUM_LAMBDA_COMPONENTS];
    long              n_set_list, n_set_atoms, n_set_params, n_set_lambdas, n_compute;
    char              err[256];
    /* device-resident entry points (the GPU-route shim, integration/gromacs_shim/fepb200_gpu_shim.h): in a
     * CPU test the caller's "device" arrays are host arrays */
    float*  x_step;      /* coordinates of the step, rvec[natoms] */
    float   sv_step[3 * FEPB200_NUM_SHIFT_VECTORS];
    int     have_x, have_result, flags_result;
    float * r_f, r_fshift[3 * FEPB200_NUM_SHIFT_VECTORS];
    double *r_vc, *r_vv, r_dvdl[2], *r_fe, *r_fdvdl;
    long    n_launch, n_set_stream;
};

static double* dup_f2d(const float* a, size_t n)
{
    double* d = (double*)malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; i < n; i++)
    {
        d[i] = a[i];
    }
    return d;
}
static int* dup_i(const int* a, size_t n) /* n is a small non-negative count */
{
    int* d = (int*)malloc(sizeof(int) * (n ? n : 1));
    if (n)
    {
        memcpy(d, a, sizeof(int) * n);
    }
    return d;
}

int fepb200_create(fepb200_ctx** ctx, int device_ordinal)
{
    (void)device_ordinal;
    *ctx = (fepb200_ctx*)calloc(1, sizeof(fepb200_ctx));
    return FEPB200_OK;
}

int fepb200_destroy(fepb200_ctx* c)
{
    free(c);
    return FEPB200_OK;
}

const char* fepb200_last_error(const fepb200_ctx* c)
{
    return c ? c->err : "";
}

const char* fepb200_describe(const fepb200_ctx* c)
{
    (void)c;
    return "fepb200 CPU STAND-IN for shim tests (oracle/fep_oracle.c), not the product";
}

int fepb200_set_params(fepb200_ctx* c, const fepb200_params* q)
{
    fep_oracle_params* p      = &c->p;
    p->eeltype                = q->eeltype;
    p->vdwtype                = q->vdwtype;
    p->vdw_modifier           = q->vdw_modifier;
    p->epsfac                 = q->epsfac;
    p->rcoulomb               = q->rcoulomb;
    p->rvdw                   = q->rvdw;
    p->rvdw_switch            = q->rvdw_switch;
    p->krf                    = q->reactionFieldCoefficient;
    p->crf                    = q->reactionFieldShift;
    p->sh_ewald               = q->sh_ewald;
    p->sh_lj_ewald            = q->sh_lj_ewald;
    p->ewaldcoeff_q           = q->ewaldcoeff_q;
    p->ewaldcoeff_lj          = q->ewaldcoeff_lj;
    p->dispersion_cpot        = q->dispersion_shift_cpot;
    p->repulsion_cpot         = q->repulsion_shift_cpot;
    p->softcoreType           = q->softcoreType;
    p->alphaVdw               = q->alphaVdw;
    p->alphaCoulomb           = q->alphaCoulomb;
    p->lambdaPower            = q->lambdaPower;
    p->sigma6WithInvalidSigma = q->sigma6WithInvalidSigma;
    p->sigma6Minimum          = q->sigma6Minimum;
    p->gapsysScaleVdW         = q->gapsysScaleLinpointVdW;
    p->gapsysScaleCoul        = q->gapsysScaleLinpointCoul;
    p->gapsysSigma6VdW        = q->gapsysSigma6VdW;
    c->have_params            = 1;
    c->n_set_params++;
    return FEPB200_OK;
}

int fepb200_set_nbfp(fepb200_ctx* c, int ntype, const float* nbfp, const float* nbfp_grid)
{
    const size_t n = 2 * (size_t)ntype * ntype;
    free(c->nbfp);
    free(c->nbfp_grid);
    c->ntype     = ntype;
    c->nbfp      = dup_f2d(nbfp, n);
    c->nbfp_grid = nbfp_grid ? dup_f2d(nbfp_grid, n) : (double*)calloc(n, sizeof(double));
    return FEPB200_OK;
}

int fepb200_set_atoms(fepb200_ctx* c, int natoms, const float* qA, const float* qB, const int* typeA, const int* typeB)
{
    free(c->qA);
    free(c->qB);
    free(c->typeA);
    free(c->typeB);
    c->natoms = natoms;
    c->qA     = dup_f2d(qA, natoms);
    c->qB     = dup_f2d(qB, natoms);
    c->typeA  = dup_i(typeA, natoms);
    c->typeB  = dup_i(typeB, natoms);
    c->n_set_atoms++;
    return FEPB200_OK;
}

int fepb200_set_list(fepb200_ctx* c, int nri, const int* iinr, const int* gid, const int* shift, const int* jindex,
                     const int* jjnr, const int* excl_fep, int nenergrp_pairs, int rank, int nranks)
{
    if (rank != 0 || nranks != 1)
    {
        snprintf(c->err, sizeof(c->err), "the stand-in holds one shard only");
        return FEPB200_ERR_UNSUPPORTED;
    }
    free(c->iinr);
    free(c->gid);
    free(c->shift);
    free(c->jindex);
    free(c->jjnr);
    free(c->excl);
    const int nrj = nri > 0 ? jindex[nri] : 0;
    c->nri        = nri;
    c->nrj        = nrj;
    c->ngrp       = nenergrp_pairs;
    c->iinr       = dup_i(iinr, nri);
    c->gid        = dup_i(gid, nri);
    c->shift      = dup_i(shift, nri);
    if (nri > 0)
    {
        c->jindex = dup_i(jindex, (size_t)nri + 1);
    }
    else
    {
        c->jindex = (int*)calloc(1, sizeof(int));
    }
    c->jjnr = dup_i(jjnr, nrj);
    if (excl_fep)
    {
        c->excl = dup_i(excl_fep, nrj);
    }
    else
    {
        c->excl = (int*)malloc(sizeof(int) * (nrj ? nrj : 1));
        for (int k = 0; k < nrj; k++)
        {
            c->excl[k] = 1;
        }
    }
    c->n_set_list++;
    return FEPB200_OK;
}

/* the per-thread lists of the reference, concatenated (and mapped) on the host: the stand-in has no device */
int fepb200_set_lists(fepb200_ctx* c, int n_lists, const fepb200_list_view* lists, const int* atom_map, int n_map,
                      int nenergrp_pairs, int rank, int nranks)
{
    long nri = 0, nrj = 0;
    for (int l = 0; l < n_lists; l++)
    {
        nri += lists[l].nri;
        nrj += lists[l].nri > 0 ? lists[l].jindex[lists[l].nri] : 0;
    }
    int* iinr   = (int*)malloc(sizeof(int) * (nri ? nri : 1));
    int* gid    = (int*)malloc(sizeof(int) * (nri ? nri : 1));
    int* shift  = (int*)malloc(sizeof(int) * (nri ? nri : 1));
    int* jindex = (int*)malloc(sizeof(int) * (nri + 1));
    int* jjnr   = (int*)malloc(sizeof(int) * (nrj ? nrj : 1));
    int* excl   = (int*)malloc(sizeof(int) * (nrj ? nrj : 1));
    long e = 0, k = 0;
    int  bad = 0;
    jindex[0] = 0;
    for (int l = 0; l < n_lists; l++)
    {
        const fepb200_list_view* v = &lists[l];
        for (int n = 0; n < v->nri; n++, e++)
        {
            int a = v->iinr[n];
            if (atom_map)
            {
                a = (a >= 0 && a < n_map) ? atom_map[a] : -1;
            }
            bad += a < 0 || a >= c->natoms;
            iinr[e]  = a;
            gid[e]   = v->gid[n];
            shift[e] = v->shift[n];
            for (int j = v->jindex[n]; j < v->jindex[n + 1]; j++, k++)
            {
                int b = v->jjnr[j];
                if (atom_map)
                {
                    b = (b >= 0 && b < n_map) ? atom_map[b] : -1;
                }
                bad += b < 0 || b >= c->natoms;
                jjnr[k] = b;
                excl[k] = v->excl_fep ? v->excl_fep[j] : 1;
            }
            jindex[e + 1] = (int)k;
        }
    }
    int rc;
    if (bad)
    {
        snprintf(c->err, sizeof(c->err), "%d atom indices of the lists are outside [0,%d)", bad, c->natoms);
        rc = FEPB200_ERR_INVALID_ARGUMENT;
    }
    else
    {
        rc = fepb200_set_list(c, (int)nri, iinr, gid, shift, jindex, jjnr, excl, nenergrp_pairs, rank, nranks);
    }
    free(iinr);
    free(gid);
    free(shift);
    free(jindex);
    free(jjnr);
    free(excl);
    return rc;
}

int fepb200_set_lambdas(fepb200_ctx* c, const float* lambda, int n_foreign, const float* all_lambda_coul,
                        const float* all_lambda_vdw)
{
    for (int i = 0; i < FEPB200_NUM_LAMBDA_COMPONENTS; i++)
    {
        c->lambda[i] = lambda[i];
    }
    free(c->all_c);
    free(c->all_v);
    c->nforeign = n_foreign;
    c->all_c    = dup_f2d(all_lambda_coul, n_foreign);
    c->all_v    = dup_f2d(all_lambda_vdw, n_foreign);
    c->n_set_lambdas++;
    return FEPB200_OK;
}

int fepb200_compute(fepb200_ctx* c, const float* x, const float* shiftvec, int flags, float* f, float* fshift, double* Vc,
                    double* Vv, double* dvdl, double* foreign_energy, double* foreign_dvdl)
{
    if (!c->have_params || !c->nbfp || !c->qA || !c->jindex)
    {
        snprintf(c->err, sizeof(c->err), "compute before params / nbfp / atoms / list were set");
        return FEPB200_ERR_STATE;
    }
    const int n  = c->natoms, L = c->nforeign;
    double*   xd = dup_f2d(x, 3 * (size_t)n);
    double*   sv = dup_f2d(shiftvec, 3 * FEPB200_NUM_SHIFT_VECTORS);
    double*   fd = (double*)calloc(3 * (size_t)n + 1, sizeof(double));
    double    fs[3 * FEPB200_NUM_SHIFT_VECTORS] = { 0 };
    double*   vc = (double*)calloc(c->ngrp + 1, sizeof(double));
    double*   vv = (double*)calloc(c->ngrp + 1, sizeof(double));
    double*   fe = (double*)calloc(L + 2, sizeof(double));
    double*   fv = (double*)calloc(2 * (size_t)(L + 2), sizeof(double));
    double    dv[2] = { 0, 0 }, secs[3] = { 0, 0, 0 };
    const int rc = fep_oracle_dispatch(&c->p, 0, 1, c->ntype, c->nbfp, c->nbfp_grid, n, xd, c->qA, c->qB, c->typeA, c->typeB,
                                       sv, c->nri, c->iinr, c->gid, c->shift, c->jindex, c->jjnr, c->excl, c->ngrp, flags,
                                       c->lambda, L, c->all_c, c->all_v, fd, fs, vc, vv, dv, fe, fv, 1, secs);
    if (rc == 0)
    {
        if ((flags & FEPB200_DO_FORCE) && f)
        {
            for (size_t i = 0; i < 3 * (size_t)n; i++)
            {
                f[i] += (float)fd[i];
            }
            if ((flags & FEPB200_DO_SHIFTFORCE) && fshift)
            {
                for (int i = 0; i < 3 * FEPB200_NUM_SHIFT_VECTORS; i++)
                {
                    fshift[i] += (float)fs[i];
                }
            }
        }
        if (flags & FEPB200_DO_POTENTIAL)
        {
            for (int g = 0; g < c->ngrp; g++)
            {
                Vc[g] += vc[g];
                Vv[g] += vv[g];
            }
        }
        dvdl[0] += dv[0];
        dvdl[1] += dv[1];
        if (flags & FEPB200_DO_FOREIGNLAMBDA)
        {
            for (int i = 0; i <= L; i++)
            {
                foreign_energy[i] += fe[i];
                foreign_dvdl[2 * i] += fv[2 * i];
                foreign_dvdl[2 * i + 1] += fv[2 * i + 1];
            }
        }
    }
    else
    {
        snprintf(c->err, sizeof(c->err), "oracle dispatch returned %d", rc);
    }
    free(xd);
    free(sv);
    free(fd);
    free(vc);
    free(vv);
    free(fe);
    free(fv);
    c->n_compute++;
    if (getenv("FEPB200_STANDIN_TRACE"))
    {
        fprintf(stderr, "standin: compute %ld set_list %ld set_atoms %ld set_params %ld set_lambdas %ld\n", c->n_compute,
                c->n_set_list, c->n_set_atoms, c->n_set_params, c->n_set_lambdas);
    }
    return rc == 0 ? FEPB200_OK : FEPB200_ERR_INVALID_ARGUMENT;
}

/* ---- device-resident entry points, on host arrays (tests/test_gpu_shim_cpu.py) ------------------------- */
int fepb200_set_stream(fepb200_ctx* c, void* stream)
{
    (void)stream;
    c->n_set_stream++;
    return FEPB200_OK;
}

int fepb200_gather_xq_device(fepb200_ctx* c, const float* d_xq, const float* shiftvec)
{
    if (!c->qA)
    {
        snprintf(c->err, sizeof(c->err), "gather before atoms were set");
        return FEPB200_ERR_STATE;
    }
    free(c->x_step);
    c->x_step = (float*)malloc(sizeof(float) * 3 * (size_t)(c->natoms ? c->natoms : 1));
    for (int i = 0; i < c->natoms; i++)
    {
        for (int d = 0; d < 3; d++)
        {
            c->x_step[3 * (size_t)i + d] = d_xq[4 * (size_t)i + d]; /* .w (the charge) is ignored */
        }
    }
    memcpy(c->sv_step, shiftvec, sizeof(c->sv_step));
    c->have_x      = 1;
    c->have_result = 0;
    return FEPB200_OK;
}

int fepb200_launch(fepb200_ctx* c, int flags, void* stream)
{
    (void)stream;
    if (!c->have_x)
    {
        snprintf(c->err, sizeof(c->err), "launch before coordinates were staged");
        return FEPB200_ERR_STATE;
    }
    const int n = c->natoms, L = c->nforeign, G = c->ngrp;
    free(c->r_f);
    free(c->r_vc);
    free(c->r_vv);
    free(c->r_fe);
    free(c->r_fdvdl);
    c->r_f     = (float*)calloc(3 * (size_t)n + 1, sizeof(float));
    c->r_vc    = (double*)calloc(G + 1, sizeof(double));
    c->r_vv    = (double*)calloc(G + 1, sizeof(double));
    c->r_fe    = (double*)calloc(L + 2, sizeof(double));
    c->r_fdvdl = (double*)calloc(2 * (size_t)(L + 2), sizeof(double));
    memset(c->r_fshift, 0, sizeof(c->r_fshift));
    c->r_dvdl[0] = c->r_dvdl[1] = 0;
    /* the host entry point with every output it can produce; flags decide what it fills */
    const int rc = fepb200_compute(c, c->x_step, c->sv_step, flags, c->r_f, c->r_fshift, c->r_vc, c->r_vv, c->r_dvdl, c->r_fe,
                                   c->r_fdvdl);
    c->have_result  = rc == FEPB200_OK;
    c->flags_result = flags;
    c->n_launch++;
    return rc;
}

int fepb200_add_forces_device(fepb200_ctx* c, float* d_f, int flags)
{
    if (!c->have_result)
    {
        snprintf(c->err, sizeof(c->err), "add_forces_device without a launch");
        return FEPB200_ERR_STATE;
    }
    (void)flags; /* FEPB200_ATOMIC_OUTPUTS: one host thread here */
    for (size_t i = 0; i < 3 * (size_t)c->natoms; i++)
    {
        d_f[i] += c->r_f[i];
    }
    return FEPB200_OK;
}

int fepb200_export_scalars_device(fepb200_ctx* c, int flags, float* eLJ, float* eElec, float* dvdlLJ, float* dvdlElec,
                                  float* eLJForeign, float* eElecForeign, float* dvdlLJForeign, float* dvdlElecForeign,
                                  float* fShift)
{
    (void)eElecForeign; /* the energy of a point goes whole through eLJForeign (include/fepb200.h) */
    if (!c->have_result)
    {
        snprintf(c->err, sizeof(c->err), "export_scalars_device without a launch");
        return FEPB200_ERR_STATE;
    }
    if (flags & FEPB200_DO_POTENTIAL)
    {
        double vc = 0, vv = 0;
        for (int g = 0; g < c->ngrp; g++)
        {
            vc += c->r_vc[g];
            vv += c->r_vv[g];
        }
        if (eElec) *eElec += (float)vc;
        if (eLJ) *eLJ += (float)vv;
    }
    if (dvdlElec) *dvdlElec += (float)c->r_dvdl[0];
    if (dvdlLJ) *dvdlLJ += (float)c->r_dvdl[1];
    if (flags & FEPB200_DO_FOREIGNLAMBDA)
    {
        for (int i = 0; i <= c->nforeign; i++)
        {
            if (eLJForeign) eLJForeign[i] += (float)c->r_fe[i];
            if (dvdlElecForeign) dvdlElecForeign[i] += (float)c->r_fdvdl[2 * i];
            if (dvdlLJForeign) dvdlLJForeign[i] += (float)c->r_fdvdl[2 * i + 1];
        }
    }
    if ((flags & FEPB200_DO_FORCE) && (flags & FEPB200_DO_SHIFTFORCE) && fShift)
    {
        for (int i = 0; i < 3 * FEPB200_NUM_SHIFT_VECTORS; i++)
        {
            fShift[i] += c->r_fshift[i];
        }
    }
    if (getenv("FEPB200_STANDIN_TRACE"))
    {
        fprintf(stderr, "standin: launch %ld set_list %ld set_atoms %ld set_params %ld set_lambdas %ld set_stream %ld\n",
                c->n_launch, c->n_set_list, c->n_set_atoms, c->n_set_params, c->n_set_lambdas, c->n_set_stream);
    }
    return FEPB200_OK;
}

/* ---- perturbed 1-4 pairs (integration/gromacs_shim/fepb200_pairs14_shim.h), answered by fep_oracle_pairs14 ---- */
int fep_oracle_pairs14(const fep_oracle_params* p, double fudgeQQ, int npairs, const int* iatoms, const double* c6A,
                       const double* c12A, const double* c6B, const double* c12B, const double* x, const double* qA,
                       const double* qB, const double* box_diag, int pbc_type, const int* gid, double lam_c, double lam_v,
                       double* f, double* fshift, double* Vc, double* Vv, double* dvdl);

struct fepb200_pairs14
{
    fepb200_ctx c; /* params only */
    double      fudge;
    int         natoms, npairs, ntypes, ngrp;
    int *       iatoms, *gid;
    double *    qA, *qB, *c6A, *c12A, *c6B, *c12B;
    long        n_set_pairs, n_compute, n_compute_foreign;
    char        err[256];
};

int fepb200_pairs14_create(fepb200_pairs14** h, int device_ordinal)
{
    (void)device_ordinal;
    *h = (fepb200_pairs14*)calloc(1, sizeof(fepb200_pairs14));
    return FEPB200_OK;
}

int fepb200_pairs14_destroy(fepb200_pairs14* h)
{
    free(h);
    return FEPB200_OK;
}

const char* fepb200_pairs14_last_error(const fepb200_pairs14* h)
{
    return h ? h->err : "";
}

int fepb200_pairs14_set_params(fepb200_pairs14* h, const fepb200_params* ic, float fudgeQQ)
{
    h->fudge = fudgeQQ;
    return fepb200_set_params(&h->c, ic);
}

int fepb200_pairs14_set_pairs(fepb200_pairs14* h, int natoms, const float* chargeA, const float* chargeB, int npairs,
                              const int* iatoms, int ntypes, const float* c6A, const float* c12A, const float* c6B,
                              const float* c12B, const int* gid, int nenergrp_pairs)
{
    free(h->iatoms);
    free(h->gid);
    free(h->qA);
    free(h->qB);
    free(h->c6A);
    free(h->c12A);
    free(h->c6B);
    free(h->c12B);
    h->natoms = natoms;
    h->npairs = npairs;
    h->ntypes = ntypes;
    h->ngrp   = nenergrp_pairs;
    h->iatoms = dup_i(iatoms, 3 * (size_t)npairs);
    if (gid)
    {
        h->gid = dup_i(gid, npairs);
    }
    else
    {
        h->gid = (int*)calloc(npairs ? npairs : 1, sizeof(int));
    }
    h->qA   = dup_f2d(chargeA, natoms);
    h->qB   = dup_f2d(chargeB, natoms);
    h->c6A  = dup_f2d(c6A, ntypes);
    h->c12A = dup_f2d(c12A, ntypes);
    h->c6B  = dup_f2d(c6B, ntypes);
    h->c12B = dup_f2d(c12B, ntypes);
    h->n_set_pairs++;
    return FEPB200_OK;
}

/* all foreign points in one call: the oracle, point after point */
int fepb200_pairs14_compute_foreign(fepb200_pairs14* h, const float* x, const float* box_diag, int pbc_type, int n_points,
                                    const float* lambda_coul, const float* lambda_vdw, double* energy, double* dvdl)
{
    if (!h->c.have_params || !h->iatoms)
    {
        snprintf(h->err, sizeof(h->err), "pairs14 compute_foreign before params / pairs were set");
        return FEPB200_ERR_STATE;
    }
    const int n  = h->natoms;
    double*   xd = dup_f2d(x, 3 * (size_t)n);
    double*   fd = (double*)calloc(3 * (size_t)n + 1, sizeof(double));
    double*   vc = (double*)calloc(h->ngrp + 1, sizeof(double));
    double*   vv = (double*)calloc(h->ngrp + 1, sizeof(double));
    double    bd[3] = { box_diag ? box_diag[0] : 0, box_diag ? box_diag[1] : 0, box_diag ? box_diag[2] : 0 };
    int       rc = 0;
    for (int i = 0; i < n_points && rc == 0; i++)
    {
        double fs[3 * FEPB200_NUM_SHIFT_VECTORS] = { 0 }, dv[2] = { 0, 0 };
        memset(vc, 0, sizeof(double) * (h->ngrp + 1));
        memset(vv, 0, sizeof(double) * (h->ngrp + 1));
        rc = fep_oracle_pairs14(&h->c.p, h->fudge, h->npairs, h->iatoms, h->c6A, h->c12A, h->c6B, h->c12B, xd, h->qA, h->qB, bd,
                                pbc_type, h->gid, lambda_coul[i], lambda_vdw[i], fd, fs, vc, vv, dv);
        energy[i] = 0.0;
        for (int g = 0; g < h->ngrp; g++)
        {
            energy[i] += vc[g] + vv[g];
        }
        dvdl[2 * i]     = dv[0];
        dvdl[2 * i + 1] = dv[1];
    }
    free(xd);
    free(fd);
    free(vc);
    free(vv);
    h->n_compute_foreign++;
    if (getenv("FEPB200_STANDIN_TRACE"))
    {
        fprintf(stderr, "standin: pairs14 compute_foreign %ld (%d points) npairs %d\n", h->n_compute_foreign, n_points, h->npairs);
    }
    if (rc != 0)
    {
        snprintf(h->err, sizeof(h->err), "oracle pairs14 returned %d", rc);
    }
    return rc == 0 ? FEPB200_OK : FEPB200_ERR_INVALID_ARGUMENT;
}

int fepb200_pairs14_compute(fepb200_pairs14* h, const float* x, const float* box_diag, int pbc_type, const float* lambda,
                            int flags, float* f, float* fshift, double* Vc14, double* Vv14, double* dvdl)
{
    if (!h->c.have_params || !h->iatoms)
    {
        snprintf(h->err, sizeof(h->err), "pairs14 compute before params / pairs were set");
        return FEPB200_ERR_STATE;
    }
    const int n  = h->natoms;
    double*   xd = dup_f2d(x, 3 * (size_t)n);
    double*   fd = (double*)calloc(3 * (size_t)n + 1, sizeof(double));
    double    fs[3 * FEPB200_NUM_SHIFT_VECTORS] = { 0 }, bd[3] = { box_diag[0], box_diag[1], box_diag[2] }, dv[2] = { 0, 0 };
    double*   vc = (double*)calloc(h->ngrp + 1, sizeof(double));
    double*   vv = (double*)calloc(h->ngrp + 1, sizeof(double));
    const int rc = fep_oracle_pairs14(&h->c.p, h->fudge, h->npairs, h->iatoms, h->c6A, h->c12A, h->c6B, h->c12B, xd, h->qA, h->qB,
                                      bd, pbc_type, h->gid, lambda[FEPB200_LAMBDA_COUL], lambda[FEPB200_LAMBDA_VDW], fd, fs, vc, vv,
                                      dv);
    if (rc == 0)
    {
        if (flags & FEPB200_DO_FORCE)
        {
            for (size_t i = 0; i < 3 * (size_t)n; i++)
            {
                f[i] += (float)fd[i];
            }
            if ((flags & FEPB200_DO_SHIFTFORCE) && fshift)
            {
                for (int i = 0; i < 3 * FEPB200_NUM_SHIFT_VECTORS; i++)
                {
                    fshift[i] += (float)fs[i];
                }
            }
        }
        if (flags & FEPB200_DO_POTENTIAL)
        {
            for (int g = 0; g < h->ngrp; g++)
            {
                Vc14[g] += vc[g];
                Vv14[g] += vv[g];
            }
        }
        dvdl[0] += dv[0];
        dvdl[1] += dv[1];
    }
    else
    {
        snprintf(h->err, sizeof(h->err), "oracle pairs14 returned %d", rc);
    }
    free(xd);
    free(fd);
    free(vc);
    free(vv);
    h->n_compute++;
    if (getenv("FEPB200_STANDIN_TRACE"))
    {
        fprintf(stderr, "standin: pairs14 compute %ld set_pairs %ld npairs %d\n", h->n_compute, h->n_set_pairs, h->npairs);
    }
    return rc == 0 ? FEPB200_OK : FEPB200_ERR_INVALID_ARGUMENT;
}
