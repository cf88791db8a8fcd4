/* Stand-in for <gsl/gsl_cblas.h> (GSL is absent from the image): the two Level-2/3 routines srcAna/Algebra.cpp calls,
 * with the standard CBLAS signatures and semantics (row-major, no transposition is all the reference uses);
 * implemented in stub/cblas_min.c as plain loops. */
#ifndef OB_STUB_GSL_CBLAS_H
#define OB_STUB_GSL_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_zgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE TransA, const int M, const int N,
                 const void *alpha, const void *A, const int lda, const void *X, const int incX, const void *beta,
                 void *Y, const int incY);
void cblas_zgemm(const enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE TransA, const enum CBLAS_TRANSPOSE TransB,
                 const int M, const int N, const int K, const void *alpha, const void *A, const int lda, const void *B,
                 const int ldb, const void *beta, void *C, const int ldc);
#ifdef __cplusplus
}
#endif
#endif
