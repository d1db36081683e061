// block.cu -- block_gmres! on device panels (src/block_gmres.jl:110-359).
//
// Data layout: every n x p block (X, B, W, V[k], ...) is a ROW-MAJOR panel in HBM (row r = p contiguous values),
// not the reference's column-major matrix: the sparse product then gathers one contiguous p-vector per nonzero and
// reads A once for all p right-hand sides, and the tall-skinny products stream both panels once.  B and X are
// transposed on the way in / out (the C ABI keeps the reference's column-major blocks).
//
// Kernels (all HBM-bound: p/8 flop per byte for the tall-skinny products, below the fp64 ridge for p <= 32):
//   spmm_rows     W = A P          p threads per row, A read once
//   panel_tn      G = V^T Q        (p x p) tile-staged, deterministic "last block finalises" reduction
//   panel_nn      Q = beta Q + alpha V S   (S p x p read from device memory: the Gram-Schmidt chain never
//                                           visits the host)
//   rows_diag     P = diag(d) V    (Jacobi M / N)
// The panel QR of the reference (LAPACK geqrf + orgqr, src/block_krylov_utils.jl:201-208) is CholQR2 on the
// device (two Gram matrices, two p x p Cholesky factorizations on the host) followed by the reconstruction of
// the Householder signs from the top p x p block of Q (Ballard et al., "Reconstructing Householder vectors
// from tall-skinny QR", 2014), so V[k] and the R factors equal LAPACK's, not only up to column signs.  A Gram
// matrix that is not numerically positive definite (rank-deficient block) falls back to Householder on the host.
// Everything p x p (the Hessenberg QR, the block back substitution) stays on the host like the reference's.
#include <cstring>

#include "solver_common.h"
#include "block.h"

namespace kb {

constexpr int kMaxBlockP = 32;
constexpr int kPanelTile = 64;      // rows of a panel staged per tile

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kBlock) transpose_kernel(int rows, int cols, const T* __restrict__ in, T* __restrict__ out) {
  // in: rows x cols row-major  ->  out: cols x rows row-major (i.e. `in` read as column-major cols x rows)
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    out[(size_t)c * rows + r] = in[i];
  }
}

template <class T>
__global__ void __launch_bounds__(kBlock) spmm_rows_kernel(Csr<T> A, int p, const T* __restrict__ X, T* __restrict__ Y) {
  const long long total = (long long)A.n * p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / p), c = (int)(i % p);
    const int kb = A.rowptr[row], ke = A.rowptr[row + 1];
    T acc = T(0);
    for (int k = kb; k < ke; k++) acc = add_rn(acc, mul_rn(A.val[k], __ldg(&X[(size_t)A.colind[k] * p + c])));
    Y[i] = acc;
  }
}

template <class T>
__global__ void __launch_bounds__(kBlock) rows_diag_kernel(long long total, int p, const T* __restrict__ d, const T* __restrict__ in,
                                                          T* __restrict__ out, int ldiv) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const T dv = d[i / p];
    out[i] = ldiv ? div_rn(in[i], dv) : mul_rn(dv, in[i]);
  }
}

// G(i,j) = sum_r V[r][i] Q[r][j], column-major p x p in `G`.  Threads are (pair, group): pair = (i,j), group g
// takes rows g, g+ngroups, ... of every staged tile; groups are combined in shared memory, CTAs through `part`.
template <class T>
__global__ void __launch_bounds__(kBlock) panel_tn_kernel(int n, int p, const T* __restrict__ V, const T* __restrict__ Q,
                                                         T* part, unsigned* ticket, T* G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Vs = reinterpret_cast<T*>(smem_raw);
  T* Qs = Vs + kPanelTile * p;
  T* red = Qs + kPanelTile * p;                   // kBlock entries
  const int pp = p * p;
  const int ngroups = pp >= kBlock ? 1 : kBlock / pp;
  const int npairs_thr = (pp + kBlock - 1) / kBlock;          // pairs per thread when pp > kBlock
  const int tid = threadIdx.x;
  const int group = pp >= kBlock ? 0 : tid / pp;
  const bool active = pp >= kBlock ? true : group < ngroups;
  T acc[4] = {T(0), T(0), T(0), T(0)};                          // pp <= 1024 -> at most 4 pairs per thread
  const int ntiles = (n + kPanelTile - 1) / kPanelTile;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * kPanelTile;
    const int rows = min(kPanelTile, n - r0);
    __syncthreads();
    for (int e = tid; e < rows * p; e += kBlock) {
      Vs[e] = V[(size_t)r0 * p + e];
      Qs[e] = Q[(size_t)r0 * p + e];
    }
    __syncthreads();
    if (active) {
      if (pp >= kBlock) {
        for (int q = 0; q < npairs_thr; q++) {
          const int pair = tid + q * kBlock;
          if (pair < pp) {
            const int i = pair % p, j = pair / p;
            T a = acc[q];
            for (int r = 0; r < rows; r++) a += Vs[r * p + i] * Qs[r * p + j];
            acc[q] = a;
          }
        }
      } else {
        const int pair = tid - group * pp;
        const int i = pair % p, j = pair / p;
        T a = acc[0];
        for (int r = group; r < rows; r += ngroups) a += Vs[r * p + i] * Qs[r * p + j];
        acc[0] = a;
      }
    }
  }
  // combine the row groups of this CTA (fixed order), then the CTAs
  __shared__ bool is_last;
  if (pp < kBlock) {
    __syncthreads();
    red[tid] = active ? acc[0] : T(0);
    __syncthreads();
    if (tid < pp) {
      T s = T(0);
      for (int g = 0; g < ngroups; g++) s += red[g * pp + tid];
      part[(size_t)blockIdx.x * pp + tid] = s;
    }
  } else {
    for (int q = 0; q < npairs_thr; q++) {
      const int pair = tid + q * kBlock;
      if (pair < pp) part[(size_t)blockIdx.x * pp + pair] = acc[q];
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int pair = tid; pair < pp; pair += kBlock) {
    T s = T(0);
    for (int b = 0; b < (int)gridDim.x; b++) s += __ldcg(&part[(size_t)b * pp + pair]);
    G[pair] = s;
  }
}

// Out[r][j] = beta * Out[r][j] + alpha * sum_i In[r][i] S(i,j);  S column-major p x p in device memory.
// In is staged per tile, so Out may alias In (the in-place Q <- Q S of the panel QR).
template <class T>
__global__ void __launch_bounds__(kBlock) panel_nn_kernel(int n, int p, T alpha, const T* In, const T* __restrict__ S, T beta, T* Out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Is = reinterpret_cast<T*>(smem_raw);
  T* Ss = Is + kPanelTile * p;
  const int tid = threadIdx.x;
  for (int e = tid; e < p * p; e += kBlock) Ss[e] = S[e];
  const int ntiles = (n + kPanelTile - 1) / kPanelTile;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * kPanelTile;
    const int rows = min(kPanelTile, n - r0);
    __syncthreads();
    for (int e = tid; e < rows * p; e += kBlock) Is[e] = In[(size_t)r0 * p + e];
    __syncthreads();
    for (int e = tid; e < rows * p; e += kBlock) {
      const int r = e / p, j = e % p;
      T acc = T(0);
      for (int i = 0; i < p; i++) acc = add_rn(acc, mul_rn(Is[r * p + i], Ss[i + j * p]));
      const size_t g = (size_t)r0 * p + e;
      Out[g] = beta == T(0) ? mul_rn(alpha, acc) : add_rn(mul_rn(beta, Out[g]), mul_rn(alpha, acc));
    }
  }
}

// ---------------------------------------------------------------------------
// host-side small dense algebra (column-major, like the reference's p x p blocks)
// ---------------------------------------------------------------------------
namespace dense {
template <class T> static T larfg(int n, T* alpha, T* x) {     // Householder reflector (LAPACK xLARFG without rescaling)
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
template <class T> static void apply_left(int m, int j, int c0, int c1, T* A, int lda, T tau, T* Cm, int ldc) {
  // C(j:m, c0:c1) <- (I - tau v v^T) C with v = [1; A(j+1:m, j)]
  for (int c = c0; c < c1; c++) {
    T w = Cm[j + (size_t)c * ldc];
    for (int i = j + 1; i < m; i++) w += A[i + (size_t)j * lda] * Cm[i + (size_t)c * ldc];
    w *= tau;
    Cm[j + (size_t)c * ldc] -= w;
    for (int i = j + 1; i < m; i++) Cm[i + (size_t)c * ldc] -= w * A[i + (size_t)j * lda];
  }
}
template <class T> static void geqr2(int m, int k, T* A, int ld, T* tau) {
  for (int j = 0; j < k && j < m; j++) {
    tau[j] = larfg(m - j, &A[j + (size_t)j * ld], &A[(j + 1 < m ? j + 1 : j) + (size_t)j * ld]);
    apply_left(m, j, j + 1, k, A, ld, tau[j], A, ld);
  }
}
template <class T> static void org2r(int m, int k, T* A, int ld, const T* tau) {
  for (int j = k - 1; j >= 0; j--) {
    apply_left(m, j, j + 1, k, A, ld, tau[j], A, ld);
    for (int i = j + 1; i < m; i++) A[i + (size_t)j * ld] = -tau[j] * A[i + (size_t)j * ld];
    A[j + (size_t)j * ld] = T(1) - tau[j];
    for (int i = 0; i < j; i++) A[i + (size_t)j * ld] = T(0);
  }
}
template <class T> static void orm2r_lt(int m, int nc, int k, T* A, int lda, const T* tau, T* Cm, int ldc) {
  for (int j = 0; j < k; j++) apply_left(m, j, 0, nc, A, lda, tau[j], Cm, ldc);
}
// householder!(Q, R, tau; compact=true) of a small m x k matrix
template <class T> static void householder_compact(int m, int k, T* Q, T* R, T* tau) {
  for (int i = 0; i < k * k; i++) R[i] = T(0);
  geqr2(m, k, Q, m, tau);
  for (int j = 0; j < k; j++) for (int i = 0; i <= j; i++) R[i + j * k] = Q[i + (size_t)j * m];
}
// G = R^T R (upper R); false when G is not numerically positive definite.  A pivot below rel * max(diag G) means
// cond(panel) beyond what CholQR2 can repair (cond^2 * eps ~ 1): the caller then takes the Householder path.
template <class T> static bool cholesky_upper(int p, const T* G, T* R) {
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
template <class T> static void inv_upper(int p, const T* R, T* X) {   // X = R^-1
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
template <class T> static void matmul(int p, const T* A, const T* B, T* Cm) {   // C = A B, all p x p
  for (int j = 0; j < p; j++)
    for (int i = 0; i < p; i++) {
      T s = 0;
      for (int k = 0; k < p; k++) s += A[i + k * p] * B[k + j * p];
      Cm[i + j * p] = s;
    }
}
// signs of the diagonal of LAPACK's Householder R relative to the positive-diagonal R, from the top p x p block W
// of the orthonormal factor: s_j = -sgn(w_jj) of the running Schur complement of W - S (sgn(0) = +1)
template <class T> static void householder_signs(int p, T* W, T* s) {
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

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
template <class T> static int panel_grid(int n) {
  const int ntiles = (n + kPanelTile - 1) / kPanelTile;
  return std::max(1, std::min(ntiles, sm_count() * 4));
}
template <class T> static void k_transpose(Ctx& c, int rows, int cols, const T* in, T* out) {
  if ((long long)rows * cols <= 0) return;
  transpose_kernel<T><<<stream_grid((long long)rows * cols, 1, 8), kBlock, 0, c.stream>>>(rows, cols, in, out);
  KB_CUDA(cudaGetLastError()); c.launches++;
}
template <class T> static void k_spmm(Ctx& c, const Csr<T>& A, int p, const T* X, T* Y) {
  if (A.n <= 0) return;
  spmm_rows_kernel<T><<<stream_grid((long long)A.n * p, 1, 8), kBlock, 0, c.stream>>>(A, p, X, Y);
  KB_CUDA(cudaGetLastError()); c.launches++;
}
template <class T> static void k_rows_diag(Ctx& c, int n, int p, const T* d, const T* in, T* out, bool ldiv) {
  rows_diag_kernel<T><<<stream_grid((long long)n * p, 1, 8), kBlock, 0, c.stream>>>((long long)n * p, p, d, in, out, ldiv ? 1 : 0);
  KB_CUDA(cudaGetLastError()); c.launches++;
}
template <class T> static void k_panel_tn(BlockWorkspace<T>& ws, const T* V, const T* Q, T* G) {
  Ctx& c = ws.ctx;
  const int p = ws.p;
  const size_t smem = sizeof(T) * ((size_t)2 * kPanelTile * p + kBlock);
  panel_tn_kernel<T><<<ws.grid, kBlock, smem, c.stream>>>(ws.n, p, V, Q, ws.part, c.tickets + 6, G);
  KB_CUDA(cudaGetLastError()); c.launches++;
}
template <class T> static void k_panel_nn(BlockWorkspace<T>& ws, T alpha, const T* In, const T* S, T beta, T* Out) {
  Ctx& c = ws.ctx;
  const int p = ws.p;
  const size_t smem = sizeof(T) * ((size_t)kPanelTile * p + (size_t)p * p);
  panel_nn_kernel<T><<<ws.grid, kBlock, smem, c.stream>>>(ws.n, p, alpha, In, S, beta, Out);
  KB_CUDA(cudaGetLastError()); c.launches++;
}
// block operator application: CSR (SpMM), diagonal, or a user block callback on host / device panels (column-major)
template <class T> static void block_apply(BlockWorkspace<T>& ws, const BlockOp<T>& op, const T* X, T* Y, bool ldiv) {
  Ctx& c = ws.ctx;
  const int n = ws.n, p = ws.p;
  switch (op.kind) {
    case BlockOp<T>::CSR: k_spmm<T>(c, *op.csr, p, X, Y); break;
    case BlockOp<T>::DIAG: k_rows_diag<T>(c, n, p, op.diag, X, Y, ldiv); break;
    case BlockOp<T>::HOST_CB: {
      // the callback sees the reference's column-major blocks in host memory (krylov.h:105-107)
      k_transpose<T>(c, n, p, X, ws.tmp);
      KB_CUDA(cudaMemcpyAsync(ws.hX, ws.tmp, sizeof(T) * (size_t)n * p, cudaMemcpyDeviceToHost, c.stream));
      c.sync();
      op.fn(ws.hX, ws.hY, p, op.userdata);
      KB_CUDA(cudaMemcpyAsync(ws.tmp, ws.hY, sizeof(T) * (size_t)n * p, cudaMemcpyHostToDevice, c.stream));
      k_transpose<T>(c, p, n, ws.tmp, Y);
      break;
    }
    case BlockOp<T>::DEV_CB: {
      k_transpose<T>(c, n, p, X, ws.tmp);
      c.sync();
      op.fn(ws.tmp, ws.tmp2, p, op.userdata);
      KB_CUDA(cudaDeviceSynchronize());
      k_transpose<T>(c, p, n, ws.tmp2, Y);
      break;
    }
    default: throw std::runtime_error("block operator missing");
  }
}

// householder!(Q, R, tau) with compact = false on an n x p device panel: Q <- orthonormal factor, Rout <- p x p R
template <class T> static void panel_qr(BlockWorkspace<T>& ws, T* Q, T* Rout) {
  Ctx& c = ws.ctx;
  const int n = ws.n, p = ws.p, pp = p * p;
  T* hG = ws.hsmall;            // pinned: [G | top]
  T* hTop = ws.hsmall + pp;
  std::vector<T> R1(pp), R2(pp), Tinv(pp), tmp(pp), sgn(p);
  bool ok = true;
  int failed_pass = -1;
  for (int pass = 0; pass < 2 && ok; pass++) {
    k_panel_tn<T>(ws, Q, Q, ws.dG);
    KB_CUDA(cudaMemcpyAsync(hG, ws.dG, sizeof(T) * pp, cudaMemcpyDeviceToHost, c.stream));
    if (pass == 1) KB_CUDA(cudaMemcpyAsync(hTop, Q, sizeof(T) * pp, cudaMemcpyDeviceToHost, c.stream));   // first p rows
    c.sync();
    std::vector<T>& R = pass == 0 ? R1 : R2;
    ok = dense::cholesky_upper<T>(p, hG, R.data());
    if (!ok) { failed_pass = pass; break; }
    dense::inv_upper<T>(p, R.data(), Tinv.data());
    if (pass == 1) {
      // top block of the final Q = (top of Q after pass 1) * R2^-1 ; rows of the row-major panel are rows of Q
      std::vector<T> W(pp);
      for (int r = 0; r < p; r++)
        for (int j = 0; j < p; j++) {
          T s = 0;
          for (int i = 0; i < p; i++) s += hTop[r * p + i] * Tinv[i + j * p];
          W[r + j * p] = s;
        }
      dense::householder_signs<T>(p, W.data(), sgn.data());
      for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) Tinv[i + j * p] *= sgn[j];      // fold S into Q <- Q R2^-1 S
    }
    KB_CUDA(cudaMemcpyAsync(ws.dS, Tinv.data(), sizeof(T) * pp, cudaMemcpyHostToDevice, c.stream));
    k_panel_nn<T>(ws, T(1), Q, ws.dS, T(0), Q);
    c.sync();                    // Tinv is a pageable host buffer reused by the next pass
  }
  if (ok) {
    dense::matmul<T>(p, R2.data(), R1.data(), tmp.data());               // R = S R2 R1
    for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) Rout[i + j * p] = i <= j ? sgn[i] * tmp[i + j * p] : T(0);
    return;
  }
  // rank-deficient block: LAPACK's algorithm on the host (column-major), then back to the device
  ws.qr_fallbacks++;
  if (n < p) throw std::runtime_error("block size exceeds the number of rows");
  std::vector<T> hq((size_t)n * p), tau(p);
  k_transpose<T>(c, n, p, Q, ws.tmp);
  KB_CUDA(cudaMemcpyAsync(hq.data(), ws.tmp, sizeof(T) * (size_t)n * p, cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  std::vector<T> Rh(pp);
  dense::householder_compact<T>(n, p, hq.data(), Rh.data(), tau.data());
  dense::org2r<T>(n, p, hq.data(), n, tau.data());
  KB_CUDA(cudaMemcpyAsync(ws.tmp, hq.data(), sizeof(T) * (size_t)n * p, cudaMemcpyHostToDevice, c.stream));
  k_transpose<T>(c, p, n, ws.tmp, Q);
  c.sync();
  if (failed_pass == 1) { dense::matmul<T>(p, Rh.data(), R1.data(), tmp.data()); Rh = tmp; }   // Q was already Q R1^-1
  for (int i = 0; i < pp; i++) Rout[i] = Rh[i];
}

// ---------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------
template <class T> static void ensure_small(BlockWorkspace<T>& ws, int nblocks) {
  // device Psi blocks and pinned slots for `nblocks` p x p blocks (+2 pinned blocks for the panel QR)
  const size_t pp = (size_t)ws.p * ws.p;
  while ((int)ws.dPsi.size() < nblocks) {
    T* d = nullptr;
    KB_CUDA(cudaMalloc(&d, sizeof(T) * pp));
    ws.dPsi.push_back(d);
  }
  const size_t need = (size_t)(nblocks + 2) * pp;
  if (need > ws.hsmall_cap) {
    ws.ctx.sync();
    if (ws.hsmall) KB_CUDA(cudaFreeHost(ws.hsmall));
    ws.hsmall_cap = 2 * need;
    KB_CUDA(cudaHostAlloc(&ws.hsmall, sizeof(T) * ws.hsmall_cap, cudaHostAllocDefault));
  }
}
template <class T> static T* slot(BlockWorkspace<T>& ws, int i) { return ws.hsmall + (size_t)(2 + i) * ws.p * ws.p; }

template <class T> BlockWorkspace<T>* block_ws_create(int m, int n, int p, int memory, int device) {
  const double t0 = now_seconds();
  if (m != n) throw std::runtime_error("System must be square");
  if (p < 1 || p > kMaxBlockP) throw std::runtime_error("block size p must be in 1..32 on the B200 path");
  BlockWorkspace<T>* ws = new BlockWorkspace<T>();
  try {
    ws->m = m; ws->n = n; ws->p = p;
    ws->ctx.init(device);
    int mem = memory > 0 ? memory : 5;                          // block_gmres.jl:99
    if (mem > n / p) mem = n / p;                               // block_krylov_workspaces.jl:138
    ws->memory = mem;
    const size_t np = (size_t)n * p, pp = (size_t)p * p;
    ws->X = dev_alloc<T>(np); ws->W = dev_alloc<T>(np);
    ws->tmp = dev_alloc<T>(np);
    for (int i = 0; i < mem; i++) ws->V.push_back(dev_alloc<T>(np));
    ws->Z.assign(mem, std::vector<T>(pp)); ws->R.assign((size_t)mem * (mem + 1) / 2, std::vector<T>(pp));
    ws->H.assign(mem, std::vector<T>(2 * pp)); ws->tau.assign(mem, std::vector<T>(p));
    ws->C.assign(pp, T(0)); ws->D.assign(2 * pp, T(0));
    ws->grid = panel_grid<T>(n);
    ws->part = dev_alloc<T>((size_t)ws->grid * pp);
    KB_CUDA(cudaMalloc(&ws->dG, sizeof(T) * pp));
    KB_CUDA(cudaMalloc(&ws->dS, sizeof(T) * pp));
    ensure_small(*ws, mem + 1);
  } catch (...) {
    block_ws_destroy(ws);
    throw;
  }
  ws->stats.allocation_timer = now_seconds() - t0;
  return ws;
}

template <class T> void block_ws_destroy(BlockWorkspace<T>* ws) {
  if (!ws) return;
  if (ws->ctx.stream) cudaStreamSynchronize(ws->ctx.stream);
  T* vecs[] = {ws->X, ws->dX, ws->W, ws->P, ws->Q, ws->Bbuf, ws->tmp, ws->tmp2, ws->part};
  for (T* v : vecs) dev_free(v);
  for (T* v : ws->V) dev_free(v);
  for (T* d : ws->dPsi) cudaFree(d);
  if (ws->dG) cudaFree(ws->dG);
  if (ws->dS) cudaFree(ws->dS);
  if (ws->hsmall) cudaFreeHost(ws->hsmall);
  if (ws->hX) cudaFreeHost(ws->hX);
  if (ws->hY) cudaFreeHost(ws->hY);
  ws->ctx.destroy();
  delete ws;
}

template <class T> static void alloc_panel_if(bool cond, BlockWorkspace<T>& ws, T*& v) {
  const double t0 = now_seconds();
  if (cond && !v) v = dev_alloc<T>((size_t)ws.n * ws.p);
  ws.stats.allocation_timer += now_seconds() - t0;
}

template <class T> void block_warm_start(BlockWorkspace<T>& ws, const T* X0_colmajor_dev) {
  alloc_panel_if(true, ws, ws.dX);
  k_transpose<T>(ws.ctx, ws.p, ws.n, X0_colmajor_dev, ws.dX);
  ws.ctx.sync();
  ws.warm_start = true;
}

template <class T> void block_get_X(BlockWorkspace<T>& ws, T* X_colmajor_dev) {
  k_transpose<T>(ws.ctx, ws.n, ws.p, ws.X, X_colmajor_dev);
  ws.ctx.sync();
}

// ===========================================================================
// block_gmres!  (src/block_gmres.jl:110-359)
// ===========================================================================
template <class T>
void block_gmres_solve(BlockWorkspace<T>& ws, const BlockOp<T>& A, const T* B_colmajor, const BlockOp<T>& M, const BlockOp<T>& N,
                       const SolveOpts& o) {
  const double start_time = now_seconds();
  Ctx& cx = ws.ctx;
  const int n = ws.n, p = ws.p;
  const int np = n * p;                        // panels are addressed as vectors of length n p by the BLAS-1 kernels
  if ((long long)n * p > 2147483647LL) throw std::runtime_error("n * p exceeds the 32-bit panel index");
  const size_t pp = (size_t)p * p;
  const bool history = o.history, ldiv = o.ldiv, restart = o.restart, reorth = o.reorthogonalization;
  if (o.verbose > 0) printf("BLOCK-GMRES: system of size %d with %d right-hand sides\n", n, p);
  const bool MisI = M.is_identity(), NisI = N.is_identity();
  alloc_panel_if(!MisI, ws, ws.Q);
  alloc_panel_if(!NisI, ws, ws.P);
  alloc_panel_if(restart, ws, ws.dX);
  alloc_panel_if(true, ws, ws.Bbuf);
  if (A.kind == BlockOp<T>::DEV_CB || M.kind == BlockOp<T>::DEV_CB || N.kind == BlockOp<T>::DEV_CB) alloc_panel_if(true, ws, ws.tmp2);
  if (A.kind == BlockOp<T>::HOST_CB || M.kind == BlockOp<T>::HOST_CB || N.kind == BlockOp<T>::HOST_CB) {
    if (!ws.hX) KB_CUDA(cudaHostAlloc(&ws.hX, sizeof(T) * (size_t)np, cudaHostAllocDefault));
    if (!ws.hY) KB_CUDA(cudaHostAlloc(&ws.hY, sizeof(T) * (size_t)np, cudaHostAllocDefault));
  }
  T *dX = ws.dX, *X = ws.X, *W = ws.W;
  std::vector<T*>& V = ws.V;
  std::vector<std::vector<T>>&Z = ws.Z, &R = ws.R, &H = ws.H, &tau = ws.tau;
  std::vector<T>&C = ws.C, &D = ws.D;
  Stats& stats = ws.stats;
  const bool warm_start = ws.warm_start;
  stats.reset();
  T* Q = MisI ? W : ws.Q;
  T* R0 = MisI ? W : ws.Q;
  T* Xr = restart ? dX : X;
  T* B = ws.Bbuf;
  const int ldd = 2 * p;
  T* D1 = D.data(); T* D2 = D.data() + p;      // D1 = D[1:p,:], D2 = D[p+1:2p,:]
  auto frob = [&](const std::vector<T>& Mx) { T s = 0; for (T v : Mx) s += v * v; return (T)std::sqrt(s); };

  k_transpose<T>(cx, p, n, B_colmajor, B);     // B as a row-major panel
  k_fill<T>(cx, np, X, T(0));
  if (warm_start) {
    block_apply(ws, A, dX, W, false);
    k_axpby<T>(cx, np, T(1), B, T(-1), W);     // W = B - W
    if (restart) k_axpy<T>(cx, np, T(1), dX, X);
  } else {
    k_copy<T>(cx, np, W, B);
  }
  if (!MisI) block_apply(ws, M, W, R0, ldiv);
  T RNorm = k_nrm2<T>(cx, np, R0);             // Frobenius norm
  if (history) stats.residuals.push_back(RNorm);
  const T eps_tol = tol_of<T>(o.atol) + tol_of<T>(o.rtol) * RNorm;
  const int mem = (int)V.size();
  int npass = 0, iter = 0, inner_iter = 0;
  const int itmax = o.itmax == 0 ? 2 * (n / p) : o.itmax;
  int inner_itmax = itmax;
  if (o.verbose > 0) printf("%5s  %5s  %7s  %5s\n", "pass", "k", "‖Rₖ‖", "timer");
  if (kdisplay(iter, o.verbose)) printf("%5d  %5d  %7.1e  %.2fs\n", npass, iter, (double)RNorm, now_seconds() - start_time);
  bool solved = RNorm <= eps_tol, tired = iter >= itmax, inner_tired = inner_iter >= inner_itmax;
  bool user_exit = false, overtimed = false;
  std::string status = "unknown";

  while (!(solved || tired || user_exit || overtimed)) {
    int nr = 0;
    // (the reference zero-fills V every cycle; every block read below is written first)
    for (auto& Psi : R) std::fill(Psi.begin(), Psi.end(), T(0));
    for (auto& blk : Z) std::fill(blk.begin(), blk.end(), T(0));
    if (restart) {
      k_fill<T>(cx, np, Xr, T(0));
      if (npass >= 1) {
        block_apply(ws, A, X, W, false);
        k_axpby<T>(cx, np, T(1), B, T(-1), W);
        if (!MisI) block_apply(ws, M, W, R0, ldiv);
      }
    }
    k_copy<T>(cx, np, V[0], R0);
    panel_qr<T>(ws, V[0], Z[0].data());         // Gamma (Z[1]) and V_1
    npass = npass + 1;
    inner_iter = 0;
    inner_tired = false;

    while (!(solved || inner_tired || user_exit || overtimed)) {
      inner_iter = inner_iter + 1;
      if (!restart && (inner_iter > mem)) {     // block_gmres.jl:231-239
        const double t0 = now_seconds();
        for (int i = 0; i < inner_iter; i++) R.push_back(std::vector<T>(pp, T(0)));
        H.push_back(std::vector<T>(2 * pp, T(0)));
        tau.push_back(std::vector<T>(p, T(0)));
        stats.allocation_timer += now_seconds() - t0;
      }
      ensure_small(ws, inner_iter + 1);
      T* Vk = V[inner_iter - 1];
      T* P = NisI ? Vk : ws.P;
      if (!NisI) block_apply(ws, N, Vk, P, ldiv);
      block_apply(ws, A, P, W, false);
      if (!MisI) block_apply(ws, M, W, Q, ldiv);
      // block modified Gram-Schmidt: the Psi blocks stay on the device between the product that makes them and the
      // update that consumes them; the host copies are fetched asynchronously and read after one sync
      for (int i = 0; i < inner_iter; i++) {
        k_panel_tn<T>(ws, V[i], Q, ws.dPsi[i]);
        k_panel_nn<T>(ws, T(-1), V[i], ws.dPsi[i], T(1), Q);
        KB_CUDA(cudaMemcpyAsync(slot(ws, i), ws.dPsi[i], sizeof(T) * pp, cudaMemcpyDeviceToHost, cx.stream));
      }
      cx.sync();
      for (int i = 0; i < inner_iter; i++) std::memcpy(R[nr + i].data(), slot(ws, i), sizeof(T) * pp);
      if (reorth) {
        for (int i = 0; i < inner_iter; i++) {
          k_panel_tn<T>(ws, V[i], Q, ws.dPsi[i]);
          k_panel_nn<T>(ws, T(-1), V[i], ws.dPsi[i], T(1), Q);
          KB_CUDA(cudaMemcpyAsync(slot(ws, i), ws.dPsi[i], sizeof(T) * pp, cudaMemcpyDeviceToHost, cx.stream));
        }
        cx.sync();
        for (int i = 0; i < inner_iter; i++) { const T* t = slot(ws, i); for (size_t k = 0; k < pp; k++) R[nr + i][k] += t[k]; }
      }
      panel_qr<T>(ws, Q, C.data());             // V_{k+1} in Q, Psi_{k+1,k} in C
      for (int i = 0; i < inner_iter - 1; i++) {  // previous reflections, block_gmres.jl:268-274
        for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) { D1[r + c * ldd] = R[nr + i][r + c * p]; D2[r + c * ldd] = R[nr + i + 1][r + c * p]; }
        dense::orm2r_lt<T>(2 * p, p, p, H[i].data(), 2 * p, tau[i].data(), D.data(), ldd);
        for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) { R[nr + i][r + c * p] = D1[r + c * ldd]; R[nr + i + 1][r + c * p] = D2[r + c * ldd]; }
      }
      std::vector<T>& Hk = H[inner_iter - 1];
      for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) { Hk[r + c * 2 * p] = R[nr + inner_iter - 1][r + c * p]; Hk[p + r + c * 2 * p] = C[r + c * p]; }
      dense::householder_compact<T>(2 * p, p, Hk.data(), R[nr + inner_iter - 1].data(), tau[inner_iter - 1].data());
      for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) { D1[r + c * ldd] = Z[inner_iter - 1][r + c * p]; D2[r + c * ldd] = T(0); }
      dense::orm2r_lt<T>(2 * p, p, p, Hk.data(), 2 * p, tau[inner_iter - 1].data(), D.data(), ldd);
      for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) Z[inner_iter - 1][r + c * p] = D1[r + c * ldd];
      for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) C[r + c * p] = D2[r + c * ldd];
      RNorm = frob(C);
      if (history) stats.residuals.push_back(RNorm);
      nr = nr + inner_iter;
      if (o.callback) { cx.sync(); stats.niter = iter + inner_iter; user_exit = o.callback(&ws, o.callback_user) != 0; }
      solved = RNorm <= eps_tol;
      inner_tired = restart ? inner_iter >= std::min(mem, inner_itmax) : inner_iter >= inner_itmax;
      overtimed = (now_seconds() - start_time) > o.timemax;
      if (kdisplay(iter + inner_iter, o.verbose)) printf("%5d  %5d  %7.1e  %.2fs\n", npass, iter + inner_iter, (double)RNorm, now_seconds() - start_time);
      if (!(solved || inner_tired || user_exit || overtimed)) {
        if (!restart && (inner_iter >= mem)) {
          const double t0 = now_seconds();
          V.push_back(dev_alloc<T>((size_t)np));
          Z.push_back(std::vector<T>(pp, T(0)));
          stats.allocation_timer += now_seconds() - t0;
        }
        k_copy<T>(cx, np, V[inner_iter], Q);
        for (int c = 0; c < p; c++) for (int r = 0; r < p; r++) Z[inner_iter][r + c * p] = D2[r + c * ldd];
      }
    }
    std::vector<std::vector<T>>& Y = Z;         // block back substitution, block_gmres.jl:316-324
    for (int i = inner_iter; i >= 1; i--) {
      int pos = nr + i - inner_iter;
      for (int j = inner_iter; j >= i + 1; j--) {
        for (int c = 0; c < p; c++)
          for (int r = 0; r < p; r++) {
            T acc = 0;
            for (int k = 0; k < p; k++) acc += R[pos - 1][r + k * p] * Y[j - 1][k + c * p];
            Y[i - 1][r + c * p] -= acc;
          }
        pos = pos - j + 1;
      }
      for (int c = 0; c < p; c++)
        for (int r = p - 1; r >= 0; r--) {
          T acc = Y[i - 1][r + c * p];
          for (int k = r + 1; k < p; k++) acc -= R[pos - 1][r + k * p] * Y[i - 1][k + c * p];
          Y[i - 1][r + c * p] = acc / R[pos - 1][r + r * p];
        }
    }
    ensure_small(ws, inner_iter + 1);
    for (int i = 0; i < inner_iter; i++) {      // X_r += V_i Y_i
      std::memcpy(slot(ws, i), Y[i].data(), sizeof(T) * pp);
      KB_CUDA(cudaMemcpyAsync(ws.dPsi[i], slot(ws, i), sizeof(T) * pp, cudaMemcpyHostToDevice, cx.stream));
      k_panel_nn<T>(ws, T(1), V[i], ws.dPsi[i], T(1), Xr);
    }
    if (!NisI) { k_copy<T>(cx, np, ws.P, Xr); block_apply(ws, N, ws.P, Xr, ldiv); }
    if (restart) k_axpy<T>(cx, np, T(1), Xr, X);
    cx.sync();
    inner_itmax = inner_itmax - inner_iter;
    iter = iter + inner_iter;
    tired = iter >= itmax;
    overtimed = (now_seconds() - start_time) > o.timemax;
  }
  if (o.verbose > 0) printf("\n");
  if (tired) status = "maximum number of iterations exceeded";
  if (solved) status = "solution good enough given atol and rtol";
  if (overtimed) status = "time limit exceeded";
  if (user_exit) status = "user-requested exit";
  if (warm_start && !restart) k_axpy<T>(cx, np, T(1), dX, X);
  ws.warm_start = false;
  cx.sync();
  stats.niter = iter; stats.solved = solved;
  stats.timer = now_seconds() - start_time;
  stats.status = status;
}

#define INST(T)                                                                                                  \
  template BlockWorkspace<T>* block_ws_create<T>(int, int, int, int, int);                                       \
  template void block_ws_destroy<T>(BlockWorkspace<T>*);                                                         \
  template void block_gmres_solve<T>(BlockWorkspace<T>&, const BlockOp<T>&, const T*, const BlockOp<T>&, const BlockOp<T>&, const SolveOpts&); \
  template void block_warm_start<T>(BlockWorkspace<T>&, const T*);                                               \
  template void block_get_X<T>(BlockWorkspace<T>&, T*);
INST(double)
INST(float)
#undef INST

}  // namespace kb
