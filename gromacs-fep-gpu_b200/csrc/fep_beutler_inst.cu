/*
 * fep_beutler_inst.cu -- the instantiations of fep_beutler_kernel for ONE (EWALD, MODE) pair, chosen with
 * -DFB_EWALD=0|1|2 (reaction field / Ewald / Ewald + LJ-PME) -DFB_MODE=0|1|2: nine objects compiled in parallel instead of one unit that takes minutes.
 */
#include "fep_beutler_kernel.cuh"

#ifndef FB_EWALD
#error "compile with -DFB_EWALD=0|1|2 -DFB_MODE=0|1|2"
#endif

FB_INST_DECL(FB_EWALD, FB_MODE)
{
    return force ? launch_size<FB_EWALD, FB_MODE, true>(ka, bs, c, stream, occ, chained)
                 : launch_size<FB_EWALD, FB_MODE, false>(ka, bs, c, stream, occ, chained);
}
