/*
 * tests/shim_standin/gpu_shim_driver.cpp -- TEST INFRASTRUCTURE for tests/test_gpu_shim_cpu.py.
 *
 * Plays the fork's GPU route against integration/gromacs_shim/fepb200_gpu_shim.h on the CPU: the
 * calls the hooks of nbnxm_gpu_fepb200.patch make, in the order and at the cadence the fork makes them
 * (gpu_init -> cuda_copy_fepparams -> per search step gpu_init_atomdata + gpu_init_feppairlist per locality ->
 * per step gpu_upload_shiftvec, gpu_clear_outputs, gpu_launch_kernel per locality), with host arrays standing in
 * for NBAtomDataGpu and GMX_FEPB200_LIB pointing at the test-only stand-in (the oracle behind the entry points).
 * What is checked is the shim: what it hands over and when, flag assembly, which buffers it lets the library
 * add into on which kind of step.
 *
 * Problem file: the format of gromacs-fep-gpu_b200/host/tests/dispatch_driver.cpp.  The list is cut into two
 * localities (first / second half of the i-entries).  Steps:
 *   0  search step; energies + virial + foreign lambdas
 *   1  plain force step (no energy, no virial): scalar buffers must stay as they are
 *   2  search step (atoms and lists handed over again), coordinates moved, lambda moved (slow growth: +0.125 on both
 *      components, through the per-step hook of do_force); energies + virial, no foreign
 * Result file, per step: float f[3N], eLJ, eElec, dvdlLJ, dvdlElec, eLJForeign[L+1], eElecForeign[L+1],
 * dvdlLJForeign[L+1], dvdlElecForeign[L+1], fShift[135].
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fepb200_gpu_shim.h"

template<typename T>
static std::vector<T> readv(FILE* f, size_t n)
{
    std::vector<T> v(n);
    if (n > 0 && fread(v.data(), sizeof(T), n, f) != n)
    {
        fprintf(stderr, "short read\n");
        exit(2);
    }
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        fprintf(stderr, "usage: %s problem.bin result.bin\n", argv[0]);
        return 2;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f)
    {
        perror("open");
        return 2;
    }
    auto      hdr = readv<int>(f, 8);
    const int n = hdr[0], t = hdr[1], nri = hdr[2], nrj = hdr[3], l = hdr[5];
    auto      prm      = readv<fepb200_params>(f, 1);
    auto      lambda   = readv<float>(f, 7);
    auto      allc     = readv<float>(f, l);
    auto      allv     = readv<float>(f, l);
    auto      nbfp     = readv<float>(f, 2 * (size_t)t * t);
    auto      nbfpGrid = readv<float>(f, 2 * (size_t)t * t);
    auto      x        = readv<float>(f, 3 * (size_t)n);
    auto      qA = readv<float>(f, n), qB = readv<float>(f, n);
    auto      tA = readv<int>(f, n), tB = readv<int>(f, n);
    auto      sv   = readv<float>(f, 135);
    auto      iinr = readv<int>(f, nri), gid = readv<int>(f, nri), shift = readv<int>(f, nri);
    auto      jindex = readv<int>(f, nri + 1), jjnr = readv<int>(f, nrj), excl = readv<int>(f, nrj);
    fclose(f);
    (void)nbfpGrid;
    (void)gid;

    /* the fork's interaction_const_t, from the parameter block of the problem */
    interaction_const_t ic;
    const fepb200_params& p     = prm[0];
    ic.eeltype                  = p.eeltype;
    ic.vdwtype                  = p.vdwtype;
    ic.vdw_modifier             = p.vdw_modifier;
    ic.epsfac                   = p.epsfac;
    ic.rcoulomb                 = p.rcoulomb;
    ic.rvdw                     = p.rvdw;
    ic.rvdw_switch              = p.rvdw_switch;
    ic.reactionFieldCoefficient = p.reactionFieldCoefficient;
    ic.reactionFieldShift       = p.reactionFieldShift;
    ic.sh_ewald                 = p.sh_ewald;
    ic.sh_lj_ewald              = p.sh_lj_ewald;
    ic.ewaldcoeff_q             = p.ewaldcoeff_q;
    ic.ewaldcoeff_lj            = p.ewaldcoeff_lj;
    ic.dispersion_shift.cpot    = p.dispersion_shift_cpot;
    ic.repulsion_shift.cpot     = p.repulsion_shift_cpot;
    auto& sc                    = *ic.softCoreParameters;
    sc.softcoreType             = p.softcoreType;
    sc.alphaVdw                 = p.alphaVdw;
    sc.alphaCoulomb             = p.alphaCoulomb;
    sc.lambdaPower              = p.lambdaPower;
    sc.sigma6WithInvalidSigma   = p.sigma6WithInvalidSigma;
    sc.sigma6Minimum            = p.sigma6Minimum;
    sc.gapsysScaleLinpointVdW   = p.gapsysScaleLinpointVdW;
    sc.gapsysScaleLinpointCoul  = p.gapsysScaleLinpointCoul;
    sc.gapsysSigma6VdW          = p.gapsysSigma6VdW;

    int        nbObject = 0; /* stands for the NbnxmGpu of the rank */
    const void* nb      = &nbObject;
    int        streams[2] = { 0, 0 }; /* stand for the cudaStream_t of the two localities */

    /* gpu_init, cuda_copy_fepparams (once, at set-up) */
    fepb200gpu::setInteractionConstants(nb, &ic);
    fepb200gpu::setLambdas(nb, lambda[FEPB200_LAMBDA_COUL], lambda[FEPB200_LAMBDA_VDW], l, allc.data(), allv.data());

    /* NBAtomDataGpu */
    std::vector<float> xq(4 * (size_t)n), force(3 * (size_t)n, 0.0F), fShift(135, 0.0F);
    float              eLJ = 0, eElec = 0, dvdlLJ = 0, dvdlElec = 0;
    std::vector<float> eLJF(l + 1, 0.0F), eElF(l + 1, 0.0F), dLJF(l + 1, 0.0F), dElF(l + 1, 0.0F);

    /* two localities: the first and the second half of the i-entries (argv[3] == "1": one locality with everything and
     * an empty second list, the rank without domain decomposition / with an empty non-local list) */
    const bool oneLocality = argc > 3 && argv[3][0] == '1';
    const int  cut         = oneLocality ? nri : nri / 2;
    auto       handOverLists = [&]() {
        for (int loc = 0; loc < 2; loc++)
        {
            const int        i0 = loc == 0 ? 0 : cut, i1 = loc == 0 ? cut : nri;
            std::vector<int> ji(i1 - i0 + 1);
            for (int i = i0; i <= i1; i++)
            {
                ji[i - i0] = jindex[i] - jindex[i0];
            }
            fepb200gpu::setList(nb, loc, i1 - i0, iinr.data() + i0, shift.data() + i0, ji.data(), jindex[i1] - jindex[i0],
                                jjnr.data() + jindex[i0], excl.data() + jindex[i0]);
        }
    };

    FILE* out = fopen(argv[2], "wb");
    if (!out)
    {
        perror("open result");
        return 2;
    }
    for (int step = 0; step < 3; step++)
    {
        const bool search = step != 1, energy = step != 1, virial = step != 1, foreign = step == 0 && l > 0;
        if (step == 2)
        {
            for (size_t i = 0; i < 3 * (size_t)n; i++)
            {
                x[i] += 0.003F * static_cast<float>((i * 2654435761U) % 7U) - 0.009F; /* the atoms moved */
            }
        }
        if (search)
        {
            fepb200gpu::setAtoms(nb, n, qA, qB, tA, tB, t, nbfp);
            handOverLists();
        }
        for (int i = 0; i < n; i++)
        {
            xq[4 * (size_t)i]     = x[3 * (size_t)i];
            xq[4 * (size_t)i + 1] = x[3 * (size_t)i + 1];
            xq[4 * (size_t)i + 2] = x[3 * (size_t)i + 2];
            xq[4 * (size_t)i + 3] = 99.0F; /* the charge slot: must be ignored */
        }
        fepb200gpu::setShiftVectors(nb, sv.data());
        /* do_force: the lambdas of this step (unchanged on steps 0 and 1) */
        const float dl = step == 2 ? 0.125F : 0.0F;
        fepb200gpu::setCurrentLambdas(nb, lambda[FEPB200_LAMBDA_COUL] + dl, lambda[FEPB200_LAMBDA_VDW] + dl);
        /* gpu_clear_outputs: forces every step, scalars on virial steps, foreign arrays every step */
        std::fill(force.begin(), force.end(), 0.0F);
        if (virial)
        {
            std::fill(fShift.begin(), fShift.end(), 0.0F);
            eLJ = eElec = dvdlLJ = dvdlElec = 0;
        }
        else
        {
            std::fill(fShift.begin(), fShift.end(), 7.0F); /* sentinels: nothing may be added on this step */
            eLJ = eElec = dvdlLJ = dvdlElec = 7.0F;
        }
        for (auto* v : { &eLJF, &eElF, &dLJF, &dElF })
        {
            std::fill(v->begin(), v->end(), 0.0F);
        }
        for (int loc = 0; loc < 2; loc++)
        {
            if ((loc == 0 ? cut : nri - cut) == 0)
            {
                continue; /* gpu_launch_kernel returns before the FEP launches of an empty list (nbnxm_cuda.cu:757-765) */
            }
            fepb200gpu::step(nb, loc, 0, &streams[loc], !oneLocality, energy, virial, foreign, xq.data(), force.data(), &eLJ, &eElec, &dvdlLJ,
                             &dvdlElec, foreign ? eLJF.data() : nullptr, foreign ? eElF.data() : nullptr,
                             foreign ? dLJF.data() : nullptr, foreign ? dElF.data() : nullptr, fShift.data());
        }
        fwrite(force.data(), sizeof(float), force.size(), out);
        for (float v : { eLJ, eElec, dvdlLJ, dvdlElec })
        {
            fwrite(&v, sizeof(float), 1, out);
        }
        for (auto* v : { &eLJF, &eElF, &dLJF, &dElF })
        {
            fwrite(v->data(), sizeof(float), v->size(), out);
        }
        fwrite(fShift.data(), sizeof(float), fShift.size(), out);
    }
    fclose(out);
    return 0;
}
