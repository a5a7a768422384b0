/*
 * dispatch_driver.cpp -- small C++ program used by tests/test_host_cpp.py: reads a flat binary
 * problem file, runs FreeEnergyDispatchGpu the way do_force() would (a dhdl step with forces,
 * virial and energies, then a plain force step), writes the results.  Format (little endian):
 *   int32 header[8] = {natoms, ntype, nri, nrj, G, L, numEnergyGroups, 0}
 *   fepb200_params, float lambda[7], float all_coul[L], float all_vdw[L],
 *   float nbfp[2TT], float nbfp_grid[2TT], float x[3N], float qA[N], float qB[N], int typeA[N], int typeB[N],
 *   float shiftvec[135], int iinr[nri], gid[nri], shift[nri], jindex[nri+1], jjnr[nrj], excl[nrj]
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "free_energy_dispatch_gpu.h"

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
    auto hdr = readv<int>(f, 8);
    const int n = hdr[0], t = hdr[1], nri = hdr[2], nrj = hdr[3], g = hdr[4], l = hdr[5], ng = hdr[6];
    auto prm      = readv<fepb200_params>(f, 1);
    auto lambda   = readv<float>(f, 7);
    auto allc     = readv<float>(f, l);
    auto allv     = readv<float>(f, l);
    auto nbfp     = readv<float>(f, 2 * (size_t)t * t);
    auto nbfpGrid = readv<float>(f, 2 * (size_t)t * t);
    auto x        = readv<float>(f, 3 * (size_t)n);
    auto qA = readv<float>(f, n), qB = readv<float>(f, n);
    auto tA = readv<int>(f, n), tB = readv<int>(f, n);
    auto sv   = readv<float>(f, 135);
    auto iinr = readv<int>(f, nri), gid = readv<int>(f, nri), shift = readv<int>(f, nri);
    auto jindex = readv<int>(f, nri + 1), jjnr = readv<int>(f, nrj), excl = readv<int>(f, nrj);
    fclose(f);
    if (ng * ng != g)
    {
        fprintf(stderr, "bad header\n");
        return 2;
    }
    try
    {
        fepb200::FreeEnergyDispatchGpu disp(ng, 0);
        disp.setInteractionConstants(prm[0]);
        disp.setNonbondedParameters(t, nbfp.data(), nbfpGrid.data());
        disp.setAtomPropertiesAB(n, qA.data(), qB.data(), tA.data(), tB.data());
        fepb200::LambdaTable fepvals;
        fepvals.n_lambda        = l;
        fepvals.all_lambda_coul = allc;
        fepvals.all_lambda_vdw  = allv;
        disp.setLambdas(lambda.data(), fepvals);
        fepb200::NbListView list{ nri, iinr.data(), gid.data(), shift.data(), jindex.data(), jjnr.data(), excl.data() };
        disp.setPairlist(list);

        /* step 1: dhdl step with forces, virial and energies */
        std::vector<float>  force(3 * (size_t)n, 0.0f), fshift(135, 0.0f);
        fepb200::EnergyData enerd(g, l);
        fepb200::StepWork   work;
        work.computeForces = work.computeVirial = work.computeEnergy = work.computeDhdl = true;
        disp.dispatchFreeEnergyKernels(x.data(), sv.data(), work, force.data(), fshift.data(), &enerd);
        /* step 2: plain MD step, forces only, into a second buffer */
        std::vector<float>  force2(3 * (size_t)n, 0.0f);
        fepb200::EnergyData enerd2(g, l);
        fepb200::StepWork   work2;
        disp.dispatchFreeEnergyKernels(x.data(), sv.data(), work2, force2.data(), nullptr, &enerd2);

        FILE* o = fopen(argv[2], "wb");
        fwrite(force.data(), sizeof(float), force.size(), o);
        fwrite(fshift.data(), sizeof(float), fshift.size(), o);
        fwrite(enerd.vCoulombSR.data(), sizeof(double), g, o);
        fwrite(enerd.vLJSR.data(), sizeof(double), g, o);
        fwrite(enerd.dvdl_lin.data(), sizeof(double), 7, o);
        fwrite(enerd.dvdl_nonlin.data(), sizeof(double), 7, o);
        fwrite(enerd.foreignEnergies.data(), sizeof(double), l + 1, o);
        fwrite(enerd.foreignDhdl.data(), sizeof(double), l + 1, o);
        fwrite(force2.data(), sizeof(float), force2.size(), o);
        fwrite(enerd2.dvdl_nonlin.data(), sizeof(double), 7, o);
        fclose(o);
        printf("%s\n", disp.describe().c_str());
    }
    catch (const fepb200::Error& e)
    {
        fprintf(stderr, "fepb200 error %d: %s\n", e.code(), e.what());
        return 1;
    }
    return 0;
}
