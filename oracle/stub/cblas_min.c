/* cblas_zgemv / cblas_zgemm for the row-major, untransposed calls of srcAna/Algebra.cpp:45,65 (reference BLAS
 * semantics: y = alpha A x + beta y, C = alpha A B + beta C), written for this repo because GSL is absent. */
#include "gsl/gsl_cblas.h"
#include <complex.h>
void cblas_zgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE TransA, const int M, const int N,
                 const void *alpha, const void *A, const int lda, const void *X, const int incX, const void *beta,
                 void *Y, const int incY) {
  const double complex al = *(const double complex *)alpha, be = *(const double complex *)beta;
  const double complex *a = (const double complex *)A, *x = (const double complex *)X;
  double complex *y = (double complex *)Y;
  (void)order;
  (void)TransA;
  for(int i = 0; i < M; ++i) {
    double complex s = 0;
    for(int j = 0; j < N; ++j)
      s += a[(long)i * lda + j] * x[(long)j * incX];
    y[(long)i * incY] = (be == 0 ? 0 : be * y[(long)i * incY]) + al * s;
  }
}
void cblas_zgemm(const enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE TransA, const enum CBLAS_TRANSPOSE TransB,
                 const int M, const int N, const int K, const void *alpha, const void *A, const int lda, const void *B,
                 const int ldb, const void *beta, void *C, const int ldc) {
  const double complex al = *(const double complex *)alpha, be = *(const double complex *)beta;
  const double complex *a = (const double complex *)A, *b = (const double complex *)B;
  double complex *c = (double complex *)C;
  (void)Order;
  (void)TransA;
  (void)TransB;
  for(int i = 0; i < M; ++i)
    for(int j = 0; j < N; ++j) {
      double complex s = 0;
      for(int k = 0; k < K; ++k)
        s += a[(long)i * lda + k] * b[(long)k * ldb + j];
      c[(long)i * ldc + j] = (be == 0 ? 0 : be * c[(long)i * ldc + j]) + al * s;
    }
}
