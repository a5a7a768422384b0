/*
 * fepb200_pairs14_shim.h -- the perturbed 1-4 pair interactions of the reference through libfepb200.so
 * (library side: fepb200_pairs14_* in include/fepb200.h).
 *
 * Included by the reference's src/gromacs/listed_forces/pairs.cpp through pairs_fepb200.patch.  With GMX_FEPB200
 * set, do_pairs_general() hands the pairs of its chunk that take the free-energy branch (bFreeEnergy,
 * pairs.cpp:618-624: an atom of the pair is perturbed or the A and B 1-4 parameters differ) to the library and
 * skips them in its own loop; everything else (unperturbed pairs, F_LJC14_Q, F_LJC_PAIRS_NB) stays with the
 * reference's tabulated evaluation.  Results go where free_energy_evaluate_single()'s go (pairs.cpp:806-821):
 * forces into the thread's rvec4 buffer, shift forces (virial flavours), Coulomb-14 / LJ-14 per energy-group
 * pair, dV/dlambda into dvdl[Coul] / dvdl[Vdw].  The same function serves the foreign-lambda evaluations of
 * calc_listed_lambda() (listed_forces.cpp), which call it with other lambdas and a scratch force buffer.
 *
 * Library handles are per calling thread (do_pairs_general runs inside the OpenMP loops of calcBondedForces and
 * calc_listed_lambda; the perturbed interactions are sorted to the end of the list, so they land in the chunks of
 * the last thread or two), one per pair list the thread sees.  A list is handed over again only when it changed
 * (domain decomposition, new chunking).
 * Cases the library does not cover are left to the reference code: triclinic boxes, screw pbc, a periodicity that
 * domain decomposition reduced (pbc_dx_aiuc then follows set_pbc_dd).
 */
#ifndef FEPB200_PAIRS14_SHIM_H
#define FEPB200_PAIRS14_SHIM_H

#include <algorithm>
#include <map>

/* what the two files that include this header (pairs.cpp, listed_forces.cpp) do not both include themselves */
#include "gromacs/math/vec.h"
#include "gromacs/tables/forcetable.h"
#include "gromacs/mdtypes/forcerec.h"
#include "gromacs/mdtypes/group.h"
#include "gromacs/mdtypes/interaction_const.h"
#include "gromacs/mdtypes/md_enums.h"
#include "gromacs/pbcutil/pbc.h"
#include "gromacs/topology/idef.h"
#include "gromacs/utility/arrayref.h"

#include "fepb200_shim_common.h"

namespace fepb200pairs
{

struct Handle
{
    fepb200_pairs14*   h = nullptr;
    bool               haveParams = false;
    fepb200_params     lastParams{};
    float              lastFudge = 0;
    int                natoms    = -1;
    std::vector<int>   key;      /* {itype, ai, aj} of the perturbed pairs the library holds */
    std::vector<float> qkey;     /* {qA[ai], qB[ai], qA[aj], qB[aj]} of those pairs as uploaded: same local indices after a
                                  * repartitioning can be other atoms */
    std::vector<int>   touched;  /* the atoms of those pairs, unique */
    std::vector<int>   gids;     /* the energy-group-pair indices of those pairs, unique */
    int                ngrp = 1; /* largest of them + 1: the size of the library's energy arrays */
    std::vector<float> fTmp;     /* rvec[natoms], zero outside a call */
    long               calls = 0, uploads = 0, foreignCalls = 0, foreignLookups = 0;
    /* all foreign lambda points of the current dH/dlambda step, evaluated in one library call (ForeignScope) */
    long                foreignGeneration = -1;
    std::vector<double> foreignE, foreignDvdl;
    ~Handle()
    {
        if (calls > 0)
        {
            std::fprintf(stderr,
                         "fepb200 pairs14 shim: %ld calls, %ld pair-list uploads, %zu perturbed 1-4 pairs in the last list; "
                         "foreign lambda: %ld library calls served %ld evaluations\n",
                         calls, uploads, key.size() / 3, foreignCalls, foreignLookups);
        }
    }
};

/* The handles of the calling thread.  A thread sees more than one pair list per step: the chunk it gets in the force
 * evaluation (calcBondedForces) and the one it gets in the foreign-lambda evaluations (calc_listed_lambda divides
 * only the perturbed interactions over the threads), so a handle is kept per list, a few per thread. */
inline Handle& handle(const std::vector<int>& key)
{
    constexpr size_t                  c_maxHandles = 4;
    static thread_local std::vector<Handle> all(c_maxHandles);
    static thread_local size_t              next = 0;
    for (Handle& hd : all)
    {
        if (hd.h != nullptr && hd.key == key)
        {
            return hd;
        }
    }
    for (Handle& hd : all)
    {
        if (hd.h == nullptr)
        {
            return hd;
        }
    }
    return all[next++ % c_maxHandles]; /* all in use with other lists: reuse them in turn */
}

inline bool enabled()
{
    static const bool on = fepb200shim::enabled() && std::getenv("GMX_FEPB200_NO_PAIRS14") == nullptr;
    return on;
}

/* The foreign-lambda evaluations of a dH/dlambda step (the loop over lambda points in ListedForces::calculate,
 * listed_forces.cpp:760-800, each calling calc_listed_lambda -> do_pairs_general for the perturbed 1-4 pairs): the hook of
 * listed_forces_fepb200.patch opens a scope around that loop and names all points; the first evaluation inside it asks
 * the library for ALL points at once (fepb200_pairs14_compute_foreign: one load of every pair instead of one call per
 * point), the others are served from that result.  One scope per rank at a time (the loop is serial). */
struct ForeignState
{
    bool               active = false;
    long               generation = 0;
    int                served = 0; /* evaluations of the perturbed 1-4 pairs seen in this scope: the next one is point `served` */
    std::vector<float> lamCoul, lamVdw;
};
inline ForeignState& foreignState()
{
    static thread_local ForeignState st;
    return st;
}
struct ForeignScope
{
    /* lambda: the current lambdas (point 0); allCoul / allVdw[numForeign]: fepvals->all_lambda[Coul | Vdw] */
    template<typename Lambda, typename All>
    ForeignScope(const Lambda& lambda, const All& allCoul, const All& allVdw, int numForeign)
    {
        if (!enabled() || std::getenv("GMX_FEPB200_PAIRS14_PER_POINT") != nullptr)
        {
            return;
        }
        ForeignState& st = foreignState();
        st.active = true;
        st.generation++;
        st.served = 0;
        st.lamCoul.assign(1, static_cast<float>(lambda[static_cast<int>(FreeEnergyPerturbationCouplingType::Coul)]));
        st.lamVdw.assign(1, static_cast<float>(lambda[static_cast<int>(FreeEnergyPerturbationCouplingType::Vdw)]));
        for (int i = 0; i < numForeign; i++)
        {
            st.lamCoul.push_back(static_cast<float>(allCoul[i]));
            st.lamVdw.push_back(static_cast<float>(allVdw[i]));
        }
    }
    ~ForeignScope() { foreignState().active = false; }
};

/* Returns true when the perturbed pairs of this chunk were computed by the library (the caller's loop then skips
 * them), false when they are left to the reference code. */
inline bool dispatch(bool                                computeVirial,
                     int                                 nbonds,
                     const t_iatom                       iatoms[],
                     const t_iparams                     iparams[],
                     const rvec                          x[],
                     rvec4                               f[],
                     rvec                                fshift[],
                     const struct t_pbc*                 pbc,
                     const real*                         lambda,
                     real*                               dvdl,
                     gmx::ArrayRef<const real>           chargeA,
                     gmx::ArrayRef<const real>           chargeB,
                     gmx::ArrayRef<const bool>           atomIsPerturbed,
                     gmx::ArrayRef<const unsigned short> cENER,
                     int                                 numEnergyGroups,
                     const t_forcerec*                   fr,
                     real*                               energygrp_elec,
                     real*                               energygrp_vdw)
{
    static_assert(sizeof(real) == sizeof(float), "the shim is for the mixed-precision build");
    /* geometry the library covers */
    int   pbcType = 0;
    float boxDiag[3] = { 0, 0, 0 };
    if (fr->bMolPBC)
    {
        if (pbc == nullptr || pbc->box[YY][XX] != 0 || pbc->box[ZZ][XX] != 0 || pbc->box[ZZ][YY] != 0)
        {
            return false;
        }
        if (pbc->pbcType == PbcType::Xyz && pbc->ndim_ePBC == 3)
        {
            pbcType = 1;
        }
        else if (pbc->pbcType == PbcType::XY && pbc->ndim_ePBC == 2)
        {
            pbcType = 2;
        }
        else
        {
            return false;
        }
        for (int d = 0; d < DIM; d++)
        {
            boxDiag[d] = pbc->box[d][d];
        }
    }
    /* the pairs of this chunk that take the free-energy branch (pairs.cpp:618-624) */
    std::vector<int> key;
    for (int i = 0; i < nbonds; i += 3)
    {
        const int  itype = iatoms[i], ai = iatoms[i + 1], aj = iatoms[i + 2];
        const auto& lj   = iparams[itype].lj14;
        if ((!atomIsPerturbed.empty() && (atomIsPerturbed[ai] || atomIsPerturbed[aj])) || lj.c6A != lj.c6B || lj.c12A != lj.c12B)
        {
            key.insert(key.end(), { itype, ai, aj });
        }
    }
    if (key.empty())
    {
        return true; /* nothing takes the branch: nothing to skip either */
    }
    /* The reference warns about and DROPS pairs beyond the range of its pair table (pairs.cpp:673-684); the library has
     * no such limit.  A chunk that holds such a pair stays on the reference code, warning included. */
    if (fr->pairsTable != nullptr)
    {
        const real range2 = fr->pairsTable->interactionRange * fr->pairsTable->interactionRange;
        for (size_t k = 0; k < key.size(); k += 3)
        {
            rvec dx;
            if (fr->bMolPBC)
            {
                pbc_dx_aiuc(pbc, x[key[k + 1]], x[key[k + 2]], dx);
            }
            else
            {
                rvec_sub(x[key[k + 1]], x[key[k + 2]], dx);
            }
            if (iprod(dx, dx) >= range2)
            {
                return false;
            }
        }
    }
    /* the charges of both states of the pairs' atoms, part of what identifies the list the library holds */
    const real*        qBk = chargeB.empty() ? chargeA.data() : chargeB.data();
    std::vector<float> qkey;
    qkey.reserve(4 * key.size() / 3);
    for (size_t k = 0; k < key.size(); k += 3)
    {
        qkey.insert(qkey.end(), { chargeA[key[k + 1]], qBk[key[k + 1]], chargeA[key[k + 2]], qBk[key[k + 2]] });
    }
    fepb200shim::loadSymbols();
    fepb200shim::Api& a  = fepb200shim::api();
    Handle&           hd = handle(key);
    if (!a.pairs14_create || !a.pairs14_compute)
    {
        gmx_fatal(FARGS, "libfepb200.so lacks the fepb200_pairs14_* entry points");
    }
    auto check = [&](int rc, const char* what) {
        if (rc != FEPB200_OK)
        {
            gmx_fatal(FARGS, "fepb200 pairs14 %s failed (%d): %s", what, rc, a.pairs14_last_error(hd.h));
        }
    };
    if (!hd.h)
    {
        const int rc = a.pairs14_create(&hd.h, fepb200shim::deviceForRank(fr->ic.get()));
        if (rc != FEPB200_OK)
        {
            gmx_fatal(FARGS, "fepb200_pairs14_create failed (%d): %s", rc, a.pairs14_last_error(nullptr));
        }
        static thread_local bool noted = false;
        if (!noted)
        {
            std::fprintf(stderr, "NOTE: perturbed 1-4 pairs are computed by libfepb200 (fepb200_pairs14_*)\n");
            noted = true;
        }
    }
    const fepb200_params p = fepb200shim::toParams(*fr->ic);
    if (!hd.haveParams || std::memcmp(&p, &hd.lastParams, sizeof(p)) != 0 || hd.lastFudge != fr->fudgeQQ)
    {
        check(a.pairs14_set_params(hd.h, &p, fr->fudgeQQ), "set_params");
        hd.lastParams = p;
        hd.lastFudge  = fr->fudgeQQ;
        hd.haveParams = true;
    }
    const int natoms = static_cast<int>(chargeA.size());
    if (key != hd.key || natoms != hd.natoms || qkey != hd.qkey)
    {
        /* the 1-4 types of these pairs, renumbered 0..k-1, with their lj14 parameters */
        std::map<int, int> localType;
        std::vector<float> c6A, c12A, c6B, c12B;
        std::vector<int>   ia(key.size()), gid(key.size() / 3);
        hd.touched.clear();
        for (size_t k = 0; k < key.size(); k += 3)
        {
            const int itype = key[k], ai = key[k + 1], aj = key[k + 2];
            auto      it    = localType.find(itype);
            if (it == localType.end())
            {
                it = localType.emplace(itype, static_cast<int>(c6A.size())).first;
                c6A.push_back(iparams[itype].lj14.c6A);
                c12A.push_back(iparams[itype].lj14.c12A);
                c6B.push_back(iparams[itype].lj14.c6B);
                c12B.push_back(iparams[itype].lj14.c12B);
            }
            ia[k]     = it->second;
            ia[k + 1] = ai;
            ia[k + 2] = aj;
            gid[k / 3] = GID(cENER[ai], cENER[aj], numEnergyGroups);
            hd.touched.push_back(ai);
            hd.touched.push_back(aj);
        }
        std::sort(hd.touched.begin(), hd.touched.end());
        hd.touched.erase(std::unique(hd.touched.begin(), hd.touched.end()), hd.touched.end());
        /* The energy arrays are indexed with GID(cENER[ai], cENER[aj], numEnergyGroups) exactly as the reference does
         * (pairs.cpp:609).  The count itself cannot be trusted for sizing: calc_one_bond passes nPerturbed in that
         * argument (listed_forces.cpp:348), harmless in the reference because it only ever indexes with the result.
         * So the library gets "largest index + 1" and only the indices that occur are added back. */
        hd.gids = gid;
        std::sort(hd.gids.begin(), hd.gids.end());
        hd.gids.erase(std::unique(hd.gids.begin(), hd.gids.end()), hd.gids.end());
        hd.ngrp = hd.gids.back() + 1;
        /* without perturbed atoms chargeB is chargeA (pairs.cpp:569-573) */
        const real* qB = chargeB.empty() ? chargeA.data() : chargeB.data();
        check(a.pairs14_set_pairs(hd.h, natoms, chargeA.data(), qB, static_cast<int>(key.size() / 3), ia.data(),
                                  static_cast<int>(c6A.size()), c6A.data(), c12A.data(), c6B.data(), c12B.data(), gid.data(),
                                  hd.ngrp),
              "set_pairs");
        hd.key    = key;
        hd.qkey   = qkey;
        hd.natoms = natoms;
        hd.fTmp.assign(3 * static_cast<size_t>(natoms), 0.0F);
        hd.uploads++;
    }
    float lam[FEPB200_NUM_LAMBDA_COMPONENTS];
    for (int i = 0; i < FEPB200_NUM_LAMBDA_COMPONENTS; i++)
    {
        lam[i] = lambda[i];
    }
    ForeignState& fs_ = foreignState();
    if (fs_.active && a.pairs14_compute_foreign != nullptr)
    {
        /* an energy-only evaluation of the loop over lambda points: all points come from ONE library call */
        const int np = static_cast<int>(fs_.lamCoul.size());
        const int i  = fs_.served++;
        if (i < np && lam[static_cast<int>(FreeEnergyPerturbationCouplingType::Coul)] == fs_.lamCoul[i]
            && lam[static_cast<int>(FreeEnergyPerturbationCouplingType::Vdw)] == fs_.lamVdw[i])
        {
            if (hd.foreignGeneration != fs_.generation)
            {
                hd.foreignE.assign(np, 0.0);
                hd.foreignDvdl.assign(2 * static_cast<size_t>(np), 0.0);
                check(a.pairs14_compute_foreign(hd.h, reinterpret_cast<const float*>(x), boxDiag, pbcType, np, fs_.lamCoul.data(),
                                                fs_.lamVdw.data(), hd.foreignE.data(), hd.foreignDvdl.data()),
                      "compute_foreign");
                hd.foreignGeneration = fs_.generation;
                hd.foreignCalls++;
            }
            /* only the sum over energy-group pairs and terms is used of a foreign evaluation (sum_epot into
             * foreign_term[F_EPOT], listed_forces.cpp:797-798): the point's Coulomb-14 + LJ-14 energy goes to one pair */
            energygrp_vdw[hd.gids.front()] += static_cast<real>(hd.foreignE[i]);
            dvdl[static_cast<int>(FreeEnergyPerturbationCouplingType::Coul)] += static_cast<real>(hd.foreignDvdl[2 * i]);
            dvdl[static_cast<int>(FreeEnergyPerturbationCouplingType::Vdw)] += static_cast<real>(hd.foreignDvdl[2 * i + 1]);
            hd.foreignLookups++;
            hd.calls++;
            return true;
        }
        /* not the evaluation the scope expected (another caller inside the loop): evaluate it the plain way */
    }
    std::vector<double> vc(hd.ngrp, 0.0), vv(hd.ngrp, 0.0);
    double              dv[2] = { 0, 0 };
    float               fs[3 * FEPB200_NUM_SHIFT_VECTORS] = { 0 };
    const int flags = FEPB200_DO_FORCE | FEPB200_DO_POTENTIAL | (computeVirial && fshift != nullptr ? FEPB200_DO_SHIFTFORCE : 0);
    check(a.pairs14_compute(hd.h, reinterpret_cast<const float*>(x), boxDiag, pbcType, lam, flags, hd.fTmp.data(), fs, vc.data(),
                            vv.data(), dv),
          "compute");
    for (const int at : hd.touched)
    {
        for (int d = 0; d < DIM; d++)
        {
            f[at][d] += hd.fTmp[3 * static_cast<size_t>(at) + d];
            hd.fTmp[3 * static_cast<size_t>(at) + d] = 0.0F;
        }
    }
    if (computeVirial && fshift != nullptr)
    {
        for (int s = 0; s < FEPB200_NUM_SHIFT_VECTORS; s++)
        {
            for (int d = 0; d < DIM; d++)
            {
                fshift[s][d] += fs[3 * s + d];
            }
        }
    }
    for (const int g : hd.gids)
    {
        energygrp_elec[g] += static_cast<real>(vc[g]);
        energygrp_vdw[g] += static_cast<real>(vv[g]);
    }
    dvdl[static_cast<int>(FreeEnergyPerturbationCouplingType::Coul)] += static_cast<real>(dv[0]);
    dvdl[static_cast<int>(FreeEnergyPerturbationCouplingType::Vdw)] += static_cast<real>(dv[1]);
    hd.calls++;
    return true;
}

} // namespace fepb200pairs

#endif
