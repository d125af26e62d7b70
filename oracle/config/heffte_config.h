/*
 * TEST INFRASTRUCTURE ONLY.  Hand-written stand-in for the header the reference's CMake would
 * generate from include/heffte_config.cmake.h (reference CMakeLists.txt:3, 5-41): version 2.4.1,
 * only the built-in `stock` CPU backend enabled (FFTW, MKL, CUDA-in-reference, MPI are absent here).
 * The vector ISA is chosen by the oracle Makefile (-DORACLE_AVX512 or -DORACLE_AVX2).
 */
#ifndef HEFFTE_CONFIG_H
#define HEFFTE_CONFIG_H

#define Heffte_VERSION_MAJOR 2
#define Heffte_VERSION_MINOR 4
#define Heffte_VERSION_PATCH 1
#define Heffte_GIT_HASH "oracle-in-place-build"

#if defined(ORACLE_AVX512)
#define Heffte_ENABLE_AVX
#define Heffte_ENABLE_AVX512
#elif defined(ORACLE_AVX2)
#define Heffte_ENABLE_AVX
#endif

#endif
