// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for Intel oneMKL's <mkl.h>, covering only the
// service / VML / BLAS / DFTI entry points the reference's hot-path TUs call.
// mkl_malloc must be free()-compatible: the reference releases some mkl_malloc blocks through
// its system allocator path (AlignedAllocation.h -> posix free) when JUCE_DSP_USE_INTEL_MKL is unset.
#pragma once
#include <cstddef>
#include <cstdlib>

#ifndef MKL_INT
#define MKL_INT int
#endif

inline void* mkl_malloc(size_t size, int alignment)
{
    void* p = nullptr;
    if (alignment < (int) sizeof(void*)) alignment = (int) sizeof(void*);
    if (posix_memalign(&p, (size_t) alignment, size ? size : 1) != 0) return nullptr;
    return p;
}
inline void mkl_free(void* p) { free(p); }
inline void* mkl_calloc(size_t num, size_t size, int alignment)
{
    void* p = mkl_malloc(num * size, alignment);
    if (p) __builtin_memset(p, 0, num * size);
    return p;
}
inline int mkl_set_num_threads_local(int) { return 0; }
inline void mkl_set_num_threads(int) {}
inline void mkl_set_dynamic(int) {}
inline void mkl_free_buffers() {}

#include "mkl_vml.h"
#include "mkl_cblas.h"
