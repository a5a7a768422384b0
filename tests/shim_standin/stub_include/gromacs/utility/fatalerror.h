/* TEST STUB (tests/test_gpu_shim_cpu.py): what the headers in integration/gromacs_shim use of the reference's
 * gromacs/utility/fatalerror.h, so that the shim headers compile in a test program without the GROMACS tree. */
#ifndef FEPB200_TEST_STUB_FATALERROR_H
#define FEPB200_TEST_STUB_FATALERROR_H
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#define FARGS 0, __FILE__, __LINE__
[[noreturn]] inline void gmx_fatal(int, const char* file, int line, const char* fmt, ...)
{
    std::va_list ap;
    va_start(ap, fmt);
    std::fprintf(stderr, "Fatal error (%s:%d): ", file, line);
    std::vfprintf(stderr, fmt, ap);
    std::fprintf(stderr, "\n");
    va_end(ap);
    std::exit(3);
}
#endif
