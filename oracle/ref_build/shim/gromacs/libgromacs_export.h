/* Hand-written stand-in for the cmake-generated export header (oracle/_ref build only). */
#ifndef LIBGROMACS_EXPORT_H
#define LIBGROMACS_EXPORT_H
#define LIBGROMACS_EXPORT
#define LIBGROMACS_NO_EXPORT
#define LIBGROMACS_DEPRECATED
#endif
