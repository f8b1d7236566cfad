// TEST INFRASTRUCTURE ONLY (oracle/): implementation of the ipp.h stand-in on top of MKL DFTI.
// DFTI prototypes/constants are declared here because the image ships the MKL symbols (inside
// torch's libtorch_cpu.so) but not the headers.  Values are the documented DFTI_CONFIG_PARAM /
// DFTI_CONFIG_VALUE enumerators of oneMKL 2024.x.
#include "ipp.h"
#include <cstdlib>
#include <cstring>
#include <new>

extern "C" {
typedef struct DFTI_DESCRIPTOR* DFTI_DESCRIPTOR_HANDLE;
long DftiCreateDescriptor_d_1d(DFTI_DESCRIPTOR_HANDLE*, int domain, long n);
long DftiSetValue(DFTI_DESCRIPTOR_HANDLE, int param, ...);
long DftiCommitDescriptor(DFTI_DESCRIPTOR_HANDLE);
long DftiComputeForward(DFTI_DESCRIPTOR_HANDLE, void*, ...);
long DftiComputeBackward(DFTI_DESCRIPTOR_HANDLE, void*, ...);
long DftiFreeDescriptor(DFTI_DESCRIPTOR_HANDLE*);
int MKL_Set_Num_Threads_Local(int);
}
namespace
{
constexpr int kDftiReal = 33;
constexpr int kDftiPlacement = 11, kDftiNotInplace = 44;
constexpr int kDftiConjugateEvenStorage = 10, kDftiComplexComplex = 39;
constexpr int kDftiBackwardScale = 5;
constexpr int kDftiThreadLimit = 27;

struct SpecImpl
{
    unsigned magic;
    int order;
    int n;
    DFTI_DESCRIPTOR_HANDLE h;
};
constexpr unsigned kMagic = 0x43505146u;
} // namespace

extern "C" {

Ipp8u* ippsMalloc_8u(int len)
{
    void* p = nullptr;
    const size_t bytes = len > 64 ? (size_t) len : 64;   // >= sizeof(SpecImpl): ippsFree peeks at the header
    if (posix_memalign(&p, 64, bytes) != 0) return nullptr;
    std::memset(p, 0, bytes);
    return static_cast<Ipp8u*>(p);
}

// The reference frees the spec block with ippsFree (FFTBackend.cpp destroyPlan); release the DFTI
// handle that lives inside it first.
void ippsFree(void* p)
{
    if (!p) return;
    auto* s = static_cast<SpecImpl*>(p);
    // Only blocks created by ippsFFTInit carry the magic; all blocks are >= 1 byte, spec blocks
    // are >= sizeof(SpecImpl).  Work buffers are sized 64 here so the read below is in-bounds.
    if (s->magic == kMagic && s->h)
    {
        DftiFreeDescriptor(&s->h);
        s->magic = 0;
    }
    free(p);
}

IppStatus ippsFFTGetSize_R_64f(int order, int flag, IppHintAlgorithm, int* sizeSpec, int* sizeInit, int* sizeWork)
{
    if (!sizeSpec || !sizeInit || !sizeWork) return ippStsNullPtrErr;
    if (order < 1 || order > 28) return ippStsFftOrderErr;
    if (flag != IPP_FFT_DIV_INV_BY_N) return ippStsFftFlagErr;
    *sizeSpec = (int) sizeof(SpecImpl) + 64;
    *sizeInit = 0;
    *sizeWork = 64;
    return ippStsNoErr;
}

IppStatus ippsFFTInit_R_64f(IppsFFTSpec_R_64f** spec, int order, int flag, IppHintAlgorithm, Ipp8u* specMem, Ipp8u*)
{
    if (!spec || !specMem) return ippStsNullPtrErr;
    if (flag != IPP_FFT_DIV_INV_BY_N) return ippStsFftFlagErr;
    auto* s = reinterpret_cast<SpecImpl*>(specMem);
    s->magic = 0;
    s->order = order;
    s->n = 1 << order;
    s->h = nullptr;
    MKL_Set_Num_Threads_Local(1);
    if (DftiCreateDescriptor_d_1d(&s->h, kDftiReal, (long) s->n) != 0) return ippStsMemAllocErr;
    DftiSetValue(s->h, kDftiPlacement, kDftiNotInplace);
    DftiSetValue(s->h, kDftiConjugateEvenStorage, kDftiComplexComplex);
    DftiSetValue(s->h, kDftiBackwardScale, 1.0 / (double) s->n);
    DftiSetValue(s->h, kDftiThreadLimit, 1);
    if (DftiCommitDescriptor(s->h) != 0)
    {
        DftiFreeDescriptor(&s->h);
        return ippStsErr;
    }
    s->magic = kMagic;
    *spec = reinterpret_cast<IppsFFTSpec_R_64f*>(s);
    return ippStsNoErr;
}

IppStatus ippsFFTFwd_RToCCS_64f(const Ipp64f* src, Ipp64f* dst, const IppsFFTSpec_R_64f* spec, Ipp8u*)
{
    if (!src || !dst || !spec) return ippStsNullPtrErr;
    auto* s = reinterpret_cast<const SpecImpl*>(spec);
    if (s->magic != kMagic) return ippStsContextMatchErr;
    return DftiComputeForward(s->h, const_cast<Ipp64f*>(src), dst) == 0 ? ippStsNoErr : ippStsErr;
}

IppStatus ippsFFTInv_CCSToR_64f(const Ipp64f* src, Ipp64f* dst, const IppsFFTSpec_R_64f* spec, Ipp8u*)
{
    if (!src || !dst || !spec) return ippStsNullPtrErr;
    auto* s = reinterpret_cast<const SpecImpl*>(spec);
    if (s->magic != kMagic) return ippStsContextMatchErr;
    return DftiComputeBackward(s->h, const_cast<Ipp64f*>(src), dst) == 0 ? ippStsNoErr : ippStsErr;
}
}
