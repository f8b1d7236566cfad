// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for Intel IPP's <ipp.h>, covering the real-FFT
// entry points the reference's FFTBackend.cpp binds (FFTBackend.cpp:33,48,130,146).
// Arithmetic is delegated to MKL DFTI (real, conjugate-even, CCE storage, backward scale 1/N),
// exported by torch's libtorch_cpu.so -- the same r2c/c2r contract as ippsFFT*_64f with
// IPP_FFT_DIV_INV_BY_N: forward unscaled -> CCS [re0,im0,...] (N/2+1 complex), inverse scaled 1/N.
// Implemented in ref_shim_impl.cpp.
#pragma once
#include <cstddef>

typedef unsigned char Ipp8u;
typedef double Ipp64f;
typedef int IppStatus;
enum
{
    ippStsNoErr = 0,
    ippStsErr = -2,
    ippStsNullPtrErr = -8,
    ippStsSizeErr = -6,
    ippStsMemAllocErr = -9,
    ippStsContextMatchErr = -13,
    ippStsFftOrderErr = -17,
    ippStsFftFlagErr = -18,
    ippStsBadArgErr = -5
};
enum { IPP_FFT_DIV_FWD_BY_N = 1, IPP_FFT_DIV_INV_BY_N = 2, IPP_FFT_DIV_BY_SQRTN = 4, IPP_FFT_NODIV_BY_ANY = 8 };
typedef enum { ippAlgHintNone, ippAlgHintFast, ippAlgHintAccurate } IppHintAlgorithm;

struct IppsFFTSpec_R_64f;

extern "C" {
Ipp8u* ippsMalloc_8u(int len);
void ippsFree(void* p);
IppStatus ippsFFTGetSize_R_64f(int order, int flag, IppHintAlgorithm hint, int* sizeSpec, int* sizeInit, int* sizeWork);
IppStatus ippsFFTInit_R_64f(IppsFFTSpec_R_64f** spec, int order, int flag, IppHintAlgorithm hint, Ipp8u* specMem, Ipp8u* initBuf);
IppStatus ippsFFTFwd_RToCCS_64f(const Ipp64f* src, Ipp64f* dst, const IppsFFTSpec_R_64f* spec, Ipp8u* work);
IppStatus ippsFFTInv_CCSToR_64f(const Ipp64f* src, Ipp64f* dst, const IppsFFTSpec_R_64f* spec, Ipp8u* work);
}
