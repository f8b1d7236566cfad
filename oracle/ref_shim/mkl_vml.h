// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <mkl_vml.h> (element-wise vector math).
#pragma once
inline void vdMul(int n, const double* a, const double* b, double* y) { for (int i = 0; i < n; ++i) y[i] = a[i] * b[i]; }
inline void vdAdd(int n, const double* a, const double* b, double* y) { for (int i = 0; i < n; ++i) y[i] = a[i] + b[i]; }
inline void vdSub(int n, const double* a, const double* b, double* y) { for (int i = 0; i < n; ++i) y[i] = a[i] - b[i]; }
inline unsigned int vmlSetMode(unsigned int) { return 0; }
