/* TEST STUB (tests/test_gpu_shim_cpu.py): the members of the reference's interaction_const_t
 * (mdtypes/interaction_const.h:111-184) that fepb200shim::toParams() reads, with plain types. */
#ifndef FEPB200_TEST_STUB_INTERACTION_CONST_H
#define FEPB200_TEST_STUB_INTERACTION_CONST_H
#include <memory>
struct shift_consts_t
{
    float c2 = 0, c3 = 0, cpot = 0;
};
struct interaction_const_t
{
    struct SoftCoreParameters
    {
        int   softcoreType = 0;
        float alphaVdw = 0, alphaCoulomb = 0;
        int   lambdaPower = 1;
        float sigma6WithInvalidSigma = 0, sigma6Minimum = 0, gapsysScaleLinpointVdW = 0, gapsysScaleLinpointCoul = 0,
              gapsysSigma6VdW = 0;
    };
    int            eeltype = 0, vdwtype = 0, vdw_modifier = 0;
    float          epsfac = 0, rcoulomb = 0, rvdw = 0, rvdw_switch = 0, reactionFieldCoefficient = 0, reactionFieldShift = 0,
                   sh_ewald = 0, sh_lj_ewald = 0, ewaldcoeff_q = 0, ewaldcoeff_lj = 0;
    shift_consts_t dispersion_shift, repulsion_shift;
    std::unique_ptr<SoftCoreParameters> softCoreParameters = std::make_unique<SoftCoreParameters>();
};
#endif
