/* Hand-written stand-in for the cmake-generated gmxpre-config.h (oracle/_ref build only). */
#define _FILE_OFFSET_BITS 64
#define GMX_FAHCORE 0
#define TMPI_WAIT_FOR_NO_ONE 0
