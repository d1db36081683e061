// dense_small.h -- host-side small dense algebra of the block path (column-major, like the reference's p x p blocks):
// the unblocked LAPACK algorithms behind householder! / kormqr! (src/block_krylov_utils.jl:192-292), the Cholesky
// pieces of CholQR2 and the Householder-sign reconstruction.  Header-only so the C ABI can export host checks
// (kb200_host_*) that run without a GPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>

namespace kb {
namespace dense {
template <class T> static inline T larfg(int n, T* alpha, T* x) {     // Householder reflector (LAPACK xLARFG without rescaling)
  if (n <= 1) return T(0);
  T xnorm = 0;
  for (int i = 0; i < n - 1; i++) xnorm += x[i] * x[i];
  xnorm = std::sqrt(xnorm);
  if (xnorm == T(0)) return T(0);
  const T a = *alpha;
  const T beta = -std::copysign(std::sqrt(a * a + xnorm * xnorm), a);
  const T tau = (beta - a) / beta;
  const T scal = T(1) / (a - beta);
  for (int i = 0; i < n - 1; i++) x[i] *= scal;
  *alpha = beta;
  return tau;
}
template <class T> static inline void apply_left(int m, int j, int c0, int c1, T* A, int lda, T tau, T* Cm, int ldc) {
  // C(j:m, c0:c1) <- (I - tau v v^T) C with v = [1; A(j+1:m, j)]
  for (int c = c0; c < c1; c++) {
    T w = Cm[j + (size_t)c * ldc];
    for (int i = j + 1; i < m; i++) w += A[i + (size_t)j * lda] * Cm[i + (size_t)c * ldc];
    w *= tau;
    Cm[j + (size_t)c * ldc] -= w;
    for (int i = j + 1; i < m; i++) Cm[i + (size_t)c * ldc] -= w * A[i + (size_t)j * lda];
  }
}
template <class T> static inline void geqr2(int m, int k, T* A, int ld, T* tau) {
  for (int j = 0; j < k && j < m; j++) {
    tau[j] = larfg(m - j, &A[j + (size_t)j * ld], &A[(j + 1 < m ? j + 1 : j) + (size_t)j * ld]);
    apply_left(m, j, j + 1, k, A, ld, tau[j], A, ld);
  }
}
template <class T> static inline void org2r(int m, int k, T* A, int ld, const T* tau) {
  for (int j = k - 1; j >= 0; j--) {
    apply_left(m, j, j + 1, k, A, ld, tau[j], A, ld);
    for (int i = j + 1; i < m; i++) A[i + (size_t)j * ld] = -tau[j] * A[i + (size_t)j * ld];
    A[j + (size_t)j * ld] = T(1) - tau[j];
    for (int i = 0; i < j; i++) A[i + (size_t)j * ld] = T(0);
  }
}
template <class T> static inline void orm2r_lt(int m, int nc, int k, T* A, int lda, const T* tau, T* Cm, int ldc) {
  for (int j = 0; j < k; j++) apply_left(m, j, 0, nc, A, lda, tau[j], Cm, ldc);
}
// householder!(Q, R, tau; compact=true) of a small m x k matrix
template <class T> static inline void householder_compact(int m, int k, T* Q, T* R, T* tau) {
  for (int i = 0; i < k * k; i++) R[i] = T(0);
  geqr2(m, k, Q, m, tau);
  for (int j = 0; j < k; j++) for (int i = 0; i <= j; i++) R[i + j * k] = Q[i + (size_t)j * m];
}
// G = R^T R (upper R); false when G is not numerically positive definite.  A pivot below rel * max(diag G) means
// cond(panel) beyond what CholQR2 can repair (cond^2 * eps ~ 1): the caller then takes the Householder path.
template <class T> static inline bool cholesky_upper(int p, const T* G, T* R) {
  const T rel = sizeof(T) == 8 ? T(1e-12) : T(1e-5);
  T gmax = 0;
  for (int j = 0; j < p; j++) gmax = std::max(gmax, G[j + j * p]);
  if (!(gmax > T(0)) || !std::isfinite(gmax)) return false;
  for (int i = 0; i < p * p; i++) R[i] = T(0);
  for (int j = 0; j < p; j++) {
    for (int i = 0; i <= j; i++) {
      T s = G[i + j * p];
      for (int k = 0; k < i; k++) s -= R[k + i * p] * R[k + j * p];
      if (i < j) R[i + j * p] = s / R[i + i * p];
      else {
        if (!(s > rel * gmax) || !std::isfinite(s)) return false;
        R[j + j * p] = std::sqrt(s);
      }
    }
  }
  return true;
}
template <class T> static inline void inv_upper(int p, const T* R, T* X) {   // X = R^-1
  for (int i = 0; i < p * p; i++) X[i] = T(0);
  for (int j = 0; j < p; j++) {
    X[j + j * p] = T(1) / R[j + j * p];
    for (int i = j - 1; i >= 0; i--) {
      T s = 0;
      for (int k = i + 1; k <= j; k++) s += R[i + k * p] * X[k + j * p];
      X[i + j * p] = -s / R[i + i * p];
    }
  }
}
template <class T> static inline void matmul(int p, const T* A, const T* B, T* Cm) {   // C = A B, all p x p
  for (int j = 0; j < p; j++)
    for (int i = 0; i < p; i++) {
      T s = 0;
      for (int k = 0; k < p; k++) s += A[i + k * p] * B[k + j * p];
      Cm[i + j * p] = s;
    }
}
// signs of the diagonal of LAPACK's Householder R relative to the positive-diagonal R, from the top p x p block W
// of the orthonormal factor: s_j = -sgn(w_jj) of the running Schur complement of W - S (sgn(0) = +1)
template <class T> static inline void householder_signs(int p, T* W, T* s) {
  for (int j = 0; j < p; j++) {
    s[j] = W[j + j * p] >= T(0) ? T(-1) : T(1);
    W[j + j * p] -= s[j];
    for (int i = j + 1; i < p; i++) {
      W[i + j * p] /= W[j + j * p];
      for (int c = j + 1; c < p; c++) W[i + c * p] -= W[i + j * p] * W[j + c * p];
    }
  }
}
}  // namespace dense

}  // namespace kb
