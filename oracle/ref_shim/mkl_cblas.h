// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <mkl_cblas.h>.
#pragma once
inline void cblas_dscal(int n, double a, double* x, int incx) { for (int i = 0; i < n; ++i) x[(long) i * incx] *= a; }
inline void cblas_dcopy(int n, const double* x, int incx, double* y, int incy) { for (int i = 0; i < n; ++i) y[(long) i * incy] = x[(long) i * incx]; }
inline void cblas_daxpy(int n, double a, const double* x, int incx, double* y, int incy) { for (int i = 0; i < n; ++i) y[(long) i * incy] += a * x[(long) i * incx]; }
