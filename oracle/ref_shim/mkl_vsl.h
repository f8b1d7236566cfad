// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for Intel oneMKL's <mkl_vsl.h>, covering the four VSL entry points
// PsychoacousticDither.h calls (:27,68-95,418-431).  oneMKL is an external, unvendored dependency of the reference
// (found via MKLROOT, version not pinned in-tree) and its SFMT19937 stream cannot be reproduced here, so the uniform
// numbers are INJECTED: a "stream" hands out, in order, the values the harness registered for it
// (cpqref_vsl_inject, ref_harness.cpp) and 0.5 once they run out.  Streams are numbered in creation order, which for
// PsychoacousticDither's constructor is the channel index (rng[i].init(...) for i = 0..MAX_CHANNELS-1, :133-140).
#pragma once
#include <cstddef>

#ifndef MKL_INT
#define MKL_INT int
#endif

struct cpqref_vsl_stream
{
    int index = 0;              // creation order since the last cpqref_vsl_begin()
    const double* values = nullptr;
    long count = 0, pos = 0;
};
typedef cpqref_vsl_stream* VSLStreamStatePtr;

#define VSL_STATUS_OK 0
#define VSL_BRNG_SFMT19937 0
#define VSL_RNG_METHOD_UNIFORM_STD 0

// defined in ref_harness.cpp
int vslNewStream(VSLStreamStatePtr* stream, int brng, unsigned int seed);
int vslDeleteStream(VSLStreamStatePtr* stream);
int vdRngUniform(int method, VSLStreamStatePtr stream, MKL_INT n, double* r, double a, double b);
