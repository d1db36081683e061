// fused_phases.cu -- fused iteration phases of bicgstab!, minres! and gmres!
// (SURVEY.md section 8a phase structures).  Each phase is ONE launch: an SpMV
// or a streaming pass whose epilogue applies the adjacent axpy/axpby/scal
// updates and accumulates the dot products the next scalar needs; the CTA that
// finishes the grid reduction derives that scalar on the device, so the phases
// of one iteration chain through device memory and the host reads the scalar
// block back once (BiCGSTAB, GMRES) or twice (MINRES) per iteration to run the
// reference's stopping logic unchanged.
//
// Arithmetic is the reference's, operation by operation (non-contracted
// mul/add in the same order as the kaxpy!/kaxpby!/kscal! sequence it replaces),
// so these paths produce the same vectors as the primitive path given the same
// scalars; tests assert that equality.
#include "kb_internal.h"
#include "spmv_tiles.cuh"

namespace kb {

// ---------------------------------------------------------------------------
// generic launchers
// ---------------------------------------------------------------------------
template <class T, int K, class Epi, class Fin, class G>
__global__ void __launch_bounds__(kTileThreads, 3) spmv_epi_tma(Csr<T> A, G xg, Epi epi, Fin fin, T* part,
                                                             unsigned* ticket, DistComm* dc) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ T sm[32];
  T d[K];
#pragma unroll
  for (int k = 0; k < K; k++) d[k] = T(0);
  spmv_tiles_run<T>(
      A, smem, xg, NoRowBegin(), [&](int row, T acc, int) { epi(row, acc, d); });
  T mine[K], tot[K];
#pragma unroll
  for (int k = 0; k < K; k++) mine[k] = block_sum(d[k], sm);
  if (grid_sum_last<T, K>(mine, part, ticket, sm, tot) && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) tot[k] = dist_reduce(dc, tot[k]);     // row-partitioned: sum over ranks
    fin(tot);
  }
}

template <class T, int K, class Epi, class Fin, class G>
__global__ void __launch_bounds__(kBlock) spmv_epi_rows(Csr<T> A, G xg, Epi epi, Fin fin, T* part,
                                                        unsigned* ticket, DistComm* dc) {
  __shared__ T sm[32];
  T d[K];
#pragma unroll
  for (int k = 0; k < K; k++) d[k] = T(0);
  const int stride = gridDim.x * blockDim.x;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < A.n; row += stride) {
    const int kb = A.rowptr[row], ke = A.rowptr[row + 1];
    T acc = T(0);
    for (int k = kb; k < ke; k++) acc = add_rn(acc, mul_rn(A.val[k], xg(A.colind[k])));
    epi(row, acc, d);
  }
  T mine[K], tot[K];
#pragma unroll
  for (int k = 0; k < K; k++) mine[k] = block_sum(d[k], sm);
  if (grid_sum_last<T, K>(mine, part, ticket, sm, tot) && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) tot[k] = dist_reduce(dc, tot[k]);
    fin(tot);
  }
}

template <class T, int K, class Body, class Fin>
__global__ void __launch_bounds__(kBlock) stream_epi(int n, Body body, Fin fin, T* part, unsigned* ticket, DistComm* dc) {
  __shared__ T sm[32];
  T d[K > 0 ? K : 1];
#pragma unroll
  for (int k = 0; k < (K > 0 ? K : 1); k++) d[k] = T(0);
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {      // two independent elements per trip
    body(i, d);
    body(i + stride, d);
  }
  if (i < n) body(i, d);
  if (K > 0) {
    T mine[K > 0 ? K : 1], tot[K > 0 ? K : 1];
#pragma unroll
    for (int k = 0; k < K; k++) mine[k] = block_sum(d[k], sm);
    if (grid_sum_last<T, (K > 0 ? K : 1)>(mine, part, ticket, sm, tot) && threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < K; k++) tot[k] = dist_reduce(dc, tot[k]);
      fin(tot);
    }
  }
}

struct NoFin {
  template <class T> __device__ void operator()(const T*) const {}
};

template <class T, int K, class Epi, class Fin, class G>
static void launch_spmv_epi_g(Ctx& c, const Csr<T>& A, G xg, Epi epi, Fin fin, int ticket) {
  if (A.tma_ok) {
    ensure_dyn_smem((const void*)spmv_epi_tma<T, K, Epi, Fin, G>, 220 * 1024);
    int occ = 0;          // persistent grid = what is really co-resident (never more than one wave)
    KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_epi_tma<T, K, Epi, Fin, G>, kTileThreads, A.smem_bytes));
    if (occ < 1) throw std::runtime_error("spmv_epi_tma does not fit on an SM with the planned shared-memory ring");
    const int grid = std::min(std::min(occ, A.ctas_per_sm) * sm_count(), std::max(1, A.ntiles));
    spmv_epi_tma<T, K, Epi, Fin, G><<<grid, kTileThreads, A.smem_bytes, c.stream>>>(A, xg, epi, fin, (T*)c.partials, c.tickets + ticket, c.dcomm);
  } else {
    spmv_epi_rows<T, K, Epi, Fin, G><<<stream_grid(A.n, 1, 8), kBlock, 0, c.stream>>>(A, xg, epi, fin, (T*)c.partials, c.tickets + ticket, c.dcomm);
  }
  KB_CUDA(cudaGetLastError());
  c.launches++;
}

template <class T, int K, class Epi, class Fin>
static void launch_spmv_epi(Ctx& c, const Csr<T>& A, const T* x, Epi epi, Fin fin, int ticket) {
  if (A.n <= 0) return;
  if (c.dex) {                               // row-partitioned operator: exchange the halo of x, gather [local | halo]
    k_halo_exchange<T>(c, x);
    launch_spmv_epi_g<T, K, Epi, Fin, XGather<T>>(c, A, xgather_of<T>(c, x), epi, fin, ticket);
  } else {
    launch_spmv_epi_g<T, K, Epi, Fin, XPlain<T>>(c, A, XPlain<T>{x}, epi, fin, ticket);
  }
}

template <class T, int K, class Body, class Fin>
static void launch_stream(Ctx& c, int n, Body body, Fin fin, int ticket) {
  if (n <= 0) return;
  stream_epi<T, K, Body, Fin><<<stream_grid(n, 2, 8), kBlock, 0, c.stream>>>(n, body, fin, (T*)c.partials, c.tickets + ticket, c.dcomm);
  KB_CUDA(cudaGetLastError());
  c.launches++;
}

template <class S> static S* state_buf(void*& dev, void*& host) {
  if (!dev) {
    KB_CUDA(cudaMalloc(&dev, kFusedBlockBytes));
    KB_CUDA(cudaMemset(dev, 0, kFusedBlockBytes));
    KB_CUDA(cudaHostAlloc(&host, kFusedBlockBytes, cudaHostAllocPortable | cudaHostAllocMapped));
  }
  static_assert(sizeof(S) <= 1024, "state block too large");
  return (S*)dev;
}

// ===========================================================================
// BiCGSTAB  (src/bicgstab.jl:215-256, N = I, M = I or a diagonal applied by multiplication)
// ===========================================================================
template <class T> struct BicgState { T rho, alpha, omega, beta, next_rho, rNorm, cv, ts, tt, cr, rr; };

template <class T> struct BicgK1Epi {   // v = M (A p) ; <c, v>          (bicgstab.jl:221-223; m: diagonal of M or null)
  T* v; const T* c; const T* m;
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const {
    if (m) acc = mul_rn(__ldg(&m[row]), acc);
    v[row] = acc; d[0] += __ldg(&c[row]) * acc;
  }
};
template <class T> struct BicgK1Fin {   // alpha = rho / <c, v>          (bicgstab.jl:223)
  BicgState<T>* s;
  __device__ void operator()(const T* tot) const { s->cv = tot[0]; s->alpha = div_rn(s->rho, tot[0]); }
};
template <class T> struct BicgK2Body {  // s = r - alpha v               (bicgstab.jl:224-225)
  const T* r; const T* v; T* sv; const BicgState<T>* s;
  __device__ __forceinline__ void operator()(int i, T*) const { sv[i] = add_rn(r[i], mul_rn(-s->alpha, v[i])); }
};
template <class T> struct BicgK3Epi {   // t = M (A s) ; <t,s>, <t,t>     (bicgstab.jl:228-230)
  T* t; const T* sv; const T* m;
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const {
    if (m) acc = mul_rn(__ldg(&m[row]), acc);
    t[row] = acc; d[0] += acc * __ldg(&sv[row]); d[1] += acc * acc;
  }
};
template <class T> struct BicgK3Fin {   // omega = <t,s>/<t,t>           (bicgstab.jl:230)
  BicgState<T>* s;
  __device__ void operator()(const T* tot) const { s->ts = tot[0]; s->tt = tot[1]; s->omega = div_rn(tot[0], tot[1]); }
};
template <class T> struct BicgK4Body {  // x += alpha p ; x += omega s ; r = s - omega t ; <c,r>, <r,r>   (:226,231-234,240)
  T* x; const T* p; const T* sv; const T* t; const T* c; T* r; const BicgState<T>* s;
  __device__ __forceinline__ void operator()(int i, T* d) const {
    const T si = sv[i];
    x[i] = add_rn(add_rn(x[i], mul_rn(s->alpha, p[i])), mul_rn(s->omega, si));
    const T rn = add_rn(si, mul_rn(-s->omega, t[i]));
    r[i] = rn;
    d[0] += c[i] * rn;
    d[1] += rn * rn;
  }
};
template <class T> struct BicgK4Fin {   // next_rho, beta = (next_rho/rho)(alpha/omega), rNorm   (:234-235,240)
  BicgState<T>* s;
  __device__ void operator()(const T* tot) const {
    s->cr = tot[0]; s->rr = tot[1];
    s->next_rho = tot[0];
    s->beta = mul_rn(div_rn(tot[0], s->rho), div_rn(s->alpha, s->omega));
    s->rNorm = sqrt_rn(tot[1]);
    s->rho = tot[0];                       // loop top of the next iteration: rho = next_rho (:218)
  }
};
template <class T> struct BicgK5Body {  // p -= omega v ; p = r + beta p   (bicgstab.jl:236-237)
  T* p; const T* r; const T* v; const BicgState<T>* s;
  __device__ __forceinline__ void operator()(int i, T*) const {
    p[i] = add_rn(r[i], mul_rn(s->beta, add_rn(p[i], mul_rn(-s->omega, v[i]))));
  }
};

// One fused BiCGSTAB iteration.  In: next_rho of the previous iteration (host value, written to the device
// block at iteration 1 only).  Out (host): alpha, omega, next_rho, rNorm -- one read-back.
template <class T>
void bicgstab_fused_iteration(Workspace<T>& ws, const Csr<T>& A, const T* cvec, bool first, T rho_in, T* alpha, T* omega,
                              T* next_rho, T* rNorm) {
  Ctx& c = ws.ctx;
  const int n = ws.n;
  typedef BicgState<T> St;
  St* S = state_buf<St>(ws.fused_state, ws.fused_host);
  St* H = (St*)ws.fused_host;
  if (first) {
    memset(H, 0, sizeof(St));
    H->rho = rho_in;
    KB_CUDA(cudaMemcpyAsync(S, H, sizeof(St), cudaMemcpyHostToDevice, c.stream));
  }
  const T* m = ws.mdiag_fused;          // left diagonal preconditioner fused into the two SpMV epilogues
  T* t = m ? ws.t : ws.qd;              // t == d == qd when M = I  (bicgstab.jl:153-154)
  launch_spmv_epi<T, 1>(c, A, ws.p, BicgK1Epi<T>{ws.v, cvec, m}, BicgK1Fin<T>{S}, 4);
  launch_stream<T, 0>(c, n, BicgK2Body<T>{ws.r, ws.v, ws.s, S}, NoFin(), 5);
  launch_spmv_epi<T, 2>(c, A, ws.s, BicgK3Epi<T>{t, ws.s, m}, BicgK3Fin<T>{S}, 4);
  launch_stream<T, 2>(c, n, BicgK4Body<T>{ws.x, ws.p, ws.s, t, cvec, ws.r, S}, BicgK4Fin<T>{S}, 5);
  KB_CUDA(cudaMemcpyAsync(H + 1, S, sizeof(St), cudaMemcpyDeviceToHost, c.stream));   // scalars of this iteration
  launch_stream<T, 0>(c, n, BicgK5Body<T>{ws.p, ws.r, ws.v, S}, NoFin(), 5);
  c.sync();
  *alpha = H[1].alpha; *omega = H[1].omega; *next_rho = H[1].next_rho; *rNorm = H[1].rNorm;
  dist_nan_guard(c, (double)*rNorm);
}

// ===========================================================================
// MINRES  (src/minres.jl:285-333,389-409, M = I or a diagonal applied by multiplication)
// ===========================================================================
template <class T> struct MinresState { T vy, alpha, beta2, xx; };

template <class T> struct MinresK1Epi {  // y = (A v [+ lambda v]) / beta [- (beta/oldbeta) r1] ; <v, y>   (:289-294)
  T* y; const T* v; const T* r1; T lambda, inv_beta, c1; int iter;     // v == r2 when M = I, else v = M r2
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const {
    const T vr = __ldg(&v[row]);
    T t = acc;
    if (lambda != T(0)) t = add_rn(t, mul_rn(lambda, vr));
    t = mul_rn(inv_beta, t);
    if (iter >= 2) t = add_rn(t, mul_rn(c1, __ldg(&r1[row])));
    y[row] = t;
    d[0] += vr * t;
  }
};
template <class T> struct MinresK1Fin {  // alpha = <v,y> / beta
  MinresState<T>* s; T beta;
  __device__ void operator()(const T* tot) const { s->vy = tot[0]; s->alpha = div_rn(tot[0], beta); }
};
template <class T> struct MinresK2Body { // y -= (alpha/beta) r2 ; w update (:295-307) ; v = M y ; <y,v> (:311-313, r2 <- y)
  T* y; const T* r2; T* w; const T* w2; const MinresState<T>* s;
  T beta, inv_beta, cs, sn, deltabar, eps; int iter;
  T* v; const T* m;                                         // M = I: v == nullptr (v aliases r2)
  __device__ __forceinline__ void operator()(int i, T* d) const {
    const T alpha = s->alpha;
    const T r2i = r2[i];
    const T vi = m ? v[i] : r2i;                            // v_k, value BEFORE v <- M y
    const T yn = add_rn(y[i], mul_rn(div_rn(-alpha, beta), r2i));
    y[i] = yn;
    if (m) {
      const T vn = mul_rn(m[i], yn);                        // v_{k+1} = M r2_{k+1}
      v[i] = vn;
      d[0] += yn * vn;
    } else {
      d[0] += yn * yn;
    }
    if (iter == 1) {
      w[i] = div_rn(vi, beta);                              // kdivcopy!(n, w, v, beta), w == w2
    } else {
      const T delta = add_rn(mul_rn(cs, deltabar), mul_rn(sn, alpha));
      T wv = w[i];                                          // w == w1
      if (iter >= 3) wv = mul_rn(-eps, wv);
      wv = add_rn(wv, mul_rn(-delta, w2[i]));
      wv = add_rn(wv, mul_rn(inv_beta, vi));
      w[i] = wv;
    }
  }
};
template <class T> struct MinresK2Fin {
  MinresState<T>* s;
  __device__ void operator()(const T* tot) const { s->beta2 = tot[0]; }
};
template <class T> struct MinresK3Body { // w /= gamma ; x += phi w ; <x,x>   (:333,389,409)
  T* w; T* x; T inv_gamma, phi;
  __device__ __forceinline__ void operator()(int i, T* d) const {
    const T wv = mul_rn(inv_gamma, w[i]);
    w[i] = wv;
    const T xn = add_rn(x[i], mul_rn(phi, wv));
    x[i] = xn;
    d[0] += xn * xn;
  }
};
template <class T> struct MinresK3Fin {
  MinresState<T>* s;
  __device__ void operator()(const T* tot) const { s->xx = tot[0]; }
};

// Phase A of a fused MINRES iteration: Lanczos step.  Rotates r1/r2/y by pointer (the reference copies).
// Returns alpha and beta_new^2 = <r2,r2>.
template <class T>
void minres_fused_lanczos(Workspace<T>& ws, const Csr<T>& A, int iter, T lambda, T beta, T oldbeta, T cs, T sn, T deltabar,
                          T eps_rot, T* w, T* alpha, T* beta2) {
  Ctx& c = ws.ctx;
  const int n = ws.n;
  typedef MinresState<T> St;
  St* S = state_buf<St>(ws.fused_state, ws.fused_host);
  St* H = (St*)ws.fused_host;
  const T inv_beta = T(1) / beta;
  const T c1 = iter >= 2 ? -beta / oldbeta : T(0);
  const T* m = ws.mdiag_fused;
  T* v = m ? ws.vv : ws.r2;                                 // minres.jl:193
  launch_spmv_epi<T, 1>(c, A, v, MinresK1Epi<T>{ws.y, v, ws.r1, lambda, inv_beta, c1, iter}, MinresK1Fin<T>{S, beta}, 4);
  launch_stream<T, 1>(c, n, MinresK2Body<T>{ws.y, ws.r2, w, ws.w2, S, beta, inv_beta, cs, sn, deltabar, eps_rot, iter,
                                            m ? ws.vv : nullptr, m},
                      MinresK2Fin<T>{S}, 5);
  KB_CUDA(cudaMemcpyAsync(H, S, sizeof(St), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  *alpha = H->alpha; *beta2 = H->beta2;
  dist_nan_guard(c, (double)*alpha + (double)*beta2);
  T* old_r1 = ws.r1;          // r1 <- r2 ; r2 <- y   (minres.jl:309-310) by rotating the bindings
  ws.r1 = ws.r2; ws.r2 = ws.y; ws.y = old_r1;
}

// Phase B: w /= gamma ; x += phi w ; returns ||x||.
template <class T>
T minres_fused_update(Workspace<T>& ws, T* w, T gamma, T phi) {
  Ctx& c = ws.ctx;
  typedef MinresState<T> St;
  St* S = state_buf<St>(ws.fused_state, ws.fused_host);
  St* H = (St*)ws.fused_host;
  launch_stream<T, 1>(c, ws.n, MinresK3Body<T>{w, ws.x, T(1) / gamma, phi}, MinresK3Fin<T>{S}, 5);
  KB_CUDA(cudaMemcpyAsync(H, S, sizeof(St), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  dist_nan_guard(c, (double)H->xx);
  return std::sqrt(H->xx);
}

// ===========================================================================
// GMRES  (src/gmres.jl:255-262,274, N = I, M = I or a diagonal applied by multiplication, no reorthogonalization)
// ===========================================================================
constexpr int kGmresMaxFused = 120;    // h[] slots in the state block
template <class T> struct GmresState { T hbis2; T h[kGmresMaxFused]; };

template <class T> struct GmresSpmvEpi {  // q = M (A v_k) ; h_1 = <v_1, q>   (gmres.jl:256-260; m: diagonal of M or null)
  T* w; const T* v1; const T* m;
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const {
    if (m) acc = mul_rn(__ldg(&m[row]), acc);
    w[row] = acc; d[0] += __ldg(&v1[row]) * acc;
  }
};
template <class T> struct GmresHFin {
  GmresState<T>* s; int slot;            // slot < 0: ||q||^2
  __device__ void operator()(const T* tot) const { if (slot >= 0) s->h[slot] = tot[0]; else s->hbis2 = tot[0]; }
};
template <class T> struct GmresMgsBody {  // q -= h_i v_i ; then <v_{i+1}, q> or <q, q>   (gmres.jl:259-262,274)
  T* q; const T* vi; const T* vnext; const GmresState<T>* s; int i;
  __device__ __forceinline__ void operator()(int j, T* d) const {
    const T qn = add_rn(q[j], mul_rn(-s->h[i], vi[j]));
    q[j] = qn;
    d[0] += (vnext ? vnext[j] : qn) * qn;
  }
};

// q = M (A xin), modified Gram-Schmidt of q against vecs[0..cnt-1] IN THAT ORDER, one launch per vector with the next
// dot product (or ||q||^2 after the last one) accumulated in the same pass; returns h[0..cnt-1] and ||q||.
// gmres! / fom! / fgmres! pass V[1..k]; dqgmres! / diom! the live window of their circular stack.
template <class T>
void fused_orth_chain(Workspace<T>& ws, const Csr<T>& A, const T* xin, T* q, const T* const* vecs, int cnt, T* h_out, T* Hbis) {
  Ctx& c = ws.ctx;
  const int n = ws.n;
  typedef GmresState<T> St;
  St* S = state_buf<St>(ws.fused_state, ws.fused_host);
  St* H = (St*)ws.fused_host;
  const T* m = ws.mdiag_fused;
  launch_spmv_epi<T, 1>(c, A, xin, GmresSpmvEpi<T>{q, vecs[0], m}, GmresHFin<T>{S, 0}, 4);
  for (int i = 0; i < cnt; i++) {
    const T* vnext = (i + 1 < cnt) ? vecs[i + 1] : nullptr;
    launch_stream<T, 1>(c, n, GmresMgsBody<T>{q, vecs[i], vnext, S, i}, GmresHFin<T>{S, (i + 1 < cnt) ? i + 1 : -1}, 5);
  }
  KB_CUDA(cudaMemcpyAsync(H, S, sizeof(T) * (size_t)(cnt + 1), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  for (int i = 0; i < cnt; i++) h_out[i] = H->h[i];
  *Hbis = std::sqrt(H->hbis2);
  dist_nan_guard(c, (double)H->hbis2);
}

// Arnoldi step k (1-based inner_iter): w = A V[k]; MGS against V[1..k]; returns h[0..k-1] and Hbis.
template <class T>
void gmres_fused_arnoldi(Workspace<T>& ws, const Csr<T>& A, int k, T* h_out, T* Hbis, const T* xin) {
  T* q = ws.mdiag_fused ? ws.q : ws.w;                      // q == w when M = I (gmres.jl:150)
  fused_orth_chain<T>(ws, A, xin ? xin : ws.V[k - 1], q, ws.V.data(), k, h_out, Hbis);
}

// xr += sum_i y_i V[i], accumulated in index order in one pass (gmres.jl:348-350)
template <class T, int NV> struct MultiAxpyBody {
  T* xr; const T* v[NV]; T y[NV]; int cnt;
  __device__ __forceinline__ void operator()(int j, T*) const {
    T acc = xr[j];
#pragma unroll
    for (int i = 0; i < NV; i++) if (i < cnt) acc = add_rn(acc, mul_rn(y[i], v[i][j]));
    xr[j] = acc;
  }
};
template <class T>
void fused_multi_axpy(Workspace<T>& ws, T* xr, int k, const T* y, T* const* vecs) {
  constexpr int NV = 8;
  for (int base = 0; base < k; base += NV) {
    MultiAxpyBody<T, NV> body;
    body.xr = xr; body.cnt = std::min(NV, k - base);
    for (int i = 0; i < NV; i++) { body.v[i] = vecs[std::min(base + i, k - 1)]; body.y[i] = (base + i < k) ? y[base + i] : T(0); }
    launch_stream<T, 0>(ws.ctx, ws.n, body, NoFin(), 5);
  }
}
template <class T>
void gmres_fused_update_x(Workspace<T>& ws, T* xr, int k, const T* y) { fused_multi_axpy<T>(ws, xr, k, y, ws.V.data()); }

// ===========================================================================
// Sibling solvers, M = N = I, CSR operator: the vector operations of one iteration are grouped into
// the fewest passes the data dependencies allow.  The scalars stay on the HOST (these loops run the reference's
// control flow unchanged and read each group of dot products back once), so unlike the four path solvers nothing
// chains through device memory; what is saved is launches, vector passes and read-backs:
//   dqgmres! / diom!  2 k + 6 launches, 2 k + 2 read-backs per iteration  ->  k + 3 launches, 1 read-back (k = window)
//   cgs!              12 launches, 3 read-backs                           ->  4 launches, 2 read-backs
//   cg_lanczos!        8 launches, 2 read-backs                           ->  3 launches, 2 read-backs
//   cr!                9 launches, 5 read-backs                           ->  3 launches, 2 read-backs
// Every element update repeats the k* sequence it replaces operation by operation (non-contracted), so vectors are
// bit-identical to the primitive path given the same scalars.
// ===========================================================================
template <class T, int K> struct StoreFin {          // K grid totals -> K consecutive device scalars
  T* out;
  __device__ void operator()(const T* tot) const {
#pragma unroll
    for (int k = 0; k < K; k++) out[k] = tot[k];
  }
};
template <class T> static T* sib_slots(Ctx& c) { return reinterpret_cast<T*>(reinterpret_cast<double*>(c.dscal) + 8); }
template <class T, int K> static void sib_read(Ctx& c, T* out) {
  T* h = reinterpret_cast<T*>(reinterpret_cast<double*>(c.hscal) + 8);
  KB_CUDA(cudaMemcpyAsync(h, sib_slots<T>(c), sizeof(T) * K, cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  for (int k = 0; k < K; k++) out[k] = h[k];
  dist_nan_guard(c, (double)out[0]);
}

// ---- dqgmres! / diom!: direction update (dqgmres.jl:279-289, diom.jl:279-289) in one pass per 8 stack vectors ----
//   for every live i: P[ppos] = -H_i P[ppos] (same slot) or P[ppos] -= H_i P[ipos];  then P[ppos] += z; P[ppos] /= H_1;
//   x += step P[ppos]
template <class T, int NV> struct TruncPBody {
  T* pp; const T* pv[NV]; T coef[NV]; int same[NV]; int cnt; int last; const T* z; T inv_h0; T step; T* x;
  __device__ __forceinline__ void operator()(int j, T*) const {
    T acc = pp[j];
#pragma unroll
    for (int i = 0; i < NV; i++)
      if (i < cnt) acc = same[i] ? mul_rn(coef[i], acc) : add_rn(acc, mul_rn(coef[i], pv[i][j]));
    if (last) {
      acc = add_rn(acc, z[j]);                    // kaxpy!(n, one, z, p)
      acc = mul_rn(inv_h0, acc);                  // kdiv!(n, p, H[1]) = kscal!(n, 1 / H[1], p)
      x[j] = add_rn(x[j], mul_rn(step, acc));
    }
    pp[j] = acc;
  }
};
template <class T>
void trunc_fused_direction(Workspace<T>& ws, T* pp, int cnt, T* const* pvecs, const T* coefs, const T* z, T h0, T step) {
  constexpr int NV = 8;
  int base = 0;
  do {
    TruncPBody<T, NV> body;
    body.pp = pp; body.cnt = std::max(0, std::min(NV, cnt - base)); body.z = z; body.inv_h0 = T(1) / h0; body.step = step; body.x = ws.x;
    body.last = (base + NV >= cnt) ? 1 : 0;
    for (int i = 0; i < NV; i++) {
      const int k = std::min(base + i, std::max(cnt - 1, 0));
      body.pv[i] = cnt > 0 ? pvecs[k] : pp; body.coef[i] = (cnt > 0 && base + i < cnt) ? coefs[base + i] : T(0);
      body.same[i] = (cnt > 0 && pvecs[k] == pp) ? 1 : 0;
    }
    launch_stream<T, 0>(ws.ctx, ws.n, body, NoFin(), 5);
    base += NV;
  } while (base < cnt);
}

// ---- cgs! (src/cgs.jl:196-239) ----
template <class T> struct CgsK1Epi {               // t = A p ; sigma = <c, t>
  T* t; const T* cv;
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const { t[row] = acc; d[0] += __ldg(&cv[row]) * acc; }
};
template <class T> struct CgsK2Body {              // q = u - alpha v ; u += q ; x += alpha u
  T* q; T* u; const T* v; T* x; T alpha;
  __device__ __forceinline__ void operator()(int j, T*) const {
    const T qn = add_rn(u[j], mul_rn(-alpha, v[j]));
    q[j] = qn;
    const T un = add_rn(u[j], qn);
    u[j] = un;
    x[j] = add_rn(x[j], mul_rn(alpha, un));
  }
};
template <class T> struct CgsK3Epi {               // s = A u ; r -= alpha s ; <c, r>, <r, r>
  T* s; T* r; const T* cv; T alpha;
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const {
    s[row] = acc;
    const T rn = add_rn(r[row], mul_rn(-alpha, acc));
    r[row] = rn;
    d[0] += __ldg(&cv[row]) * rn; d[1] += rn * rn;
  }
};
template <class T> struct CgsK4Body {              // u = r + beta q ; p = u + beta (q + beta p)
  T* u; T* p; const T* r; const T* q; T beta;
  __device__ __forceinline__ void operator()(int j, T*) const {
    const T un = add_rn(r[j], mul_rn(beta, q[j]));
    u[j] = un;
    const T p1 = add_rn(q[j], mul_rn(beta, p[j]));
    p[j] = add_rn(un, mul_rn(beta, p1));
  }
};
template <class T> T cgs_fused_sigma(Workspace<T>& ws, const Csr<T>& A, const T* cvec) {
  Ctx& c = ws.ctx;
  launch_spmv_epi<T, 1>(c, A, ws.p, CgsK1Epi<T>{ws.ts, cvec}, StoreFin<T, 1>{sib_slots<T>(c)}, 4);
  T out[1]; sib_read<T, 1>(c, out);
  return out[0];
}
template <class T> void cgs_fused_update(Workspace<T>& ws, const Csr<T>& A, const T* cvec, T alpha, T* rho_next, T* rr) {
  Ctx& c = ws.ctx;
  launch_stream<T, 0>(c, ws.n, CgsK2Body<T>{ws.q, ws.u, ws.ts, ws.x, alpha}, NoFin(), 5);
  launch_spmv_epi<T, 2>(c, A, ws.u, CgsK3Epi<T>{ws.ts, ws.r, cvec, alpha}, StoreFin<T, 2>{sib_slots<T>(c)}, 4);
  T out[2]; sib_read<T, 2>(c, out);
  *rho_next = out[0]; *rr = out[1];
}
template <class T> void cgs_fused_directions(Workspace<T>& ws, T beta) {
  launch_stream<T, 0>(ws.ctx, ws.n, CgsK4Body<T>{ws.u, ws.p, ws.r, ws.q, beta}, NoFin(), 5);
}

// ---- cg_lanczos! (src/cg_lanczos.jl:186-216), M = I so v === Mv ----
template <class T> struct LanK1Epi {               // Mv_next = A v ; delta = <v, Mv_next>
  T* mvn; const T* v;
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const { mvn[row] = acc; d[0] += __ldg(&v[row]) * acc; }
};
template <class T> struct LanK2Body {              // Mv_next -= delta Mv (- beta Mv_prev) ; Mv_prev = Mv ; Mv = Mv_next ; ||Mv||^2
  T* mvn; T* mv; T* mvp; T delta; T beta; int later;
  __device__ __forceinline__ void operator()(int j, T* d) const {
    T m = add_rn(mvn[j], mul_rn(-delta, mv[j]));
    if (later) { m = add_rn(m, mul_rn(-beta, mvp[j])); mvp[j] = mv[j]; }
    mvn[j] = m; mv[j] = m;
    d[0] += m * m;
  }
};
template <class T> struct LanK3Body {              // v /= beta ; x += gamma p ; p = sigma v + omega p
  T* v; T* x; T* p; T inv_beta; T gamma; T sigma; T omega;
  __device__ __forceinline__ void operator()(int j, T*) const {
    const T vn = mul_rn(inv_beta, v[j]);
    v[j] = vn;
    x[j] = add_rn(x[j], mul_rn(gamma, p[j]));
    p[j] = add_rn(mul_rn(sigma, vn), mul_rn(omega, p[j]));
  }
};
template <class T> T lanczos_fused_delta(Workspace<T>& ws, const Csr<T>& A) {
  Ctx& c = ws.ctx;
  launch_spmv_epi<T, 1>(c, A, ws.Mv, LanK1Epi<T>{ws.Mv_next, ws.Mv}, StoreFin<T, 1>{sib_slots<T>(c)}, 4);
  T out[1]; sib_read<T, 1>(c, out);
  return out[0];
}
template <class T> T lanczos_fused_recur(Workspace<T>& ws, T delta, T beta, bool later) {
  Ctx& c = ws.ctx;
  launch_stream<T, 1>(c, ws.n, LanK2Body<T>{ws.Mv_next, ws.Mv, ws.Mv_prev, delta, beta, later ? 1 : 0}, StoreFin<T, 1>{sib_slots<T>(c)}, 5);
  T out[1]; sib_read<T, 1>(c, out);
  return std::sqrt(out[0]);
}
template <class T> void lanczos_fused_update(Workspace<T>& ws, T beta, T gamma, T sigma, T omega) {
  launch_stream<T, 0>(ws.ctx, ws.n, LanK3Body<T>{ws.Mv, ws.x, ws.p, T(1) / beta, gamma, sigma, omega}, NoFin(), 5);
}

// ---- cr! (src/cr.jl:375-445), M = I, no trust region, no linesearch ----
template <class T> struct CrK1Body {               // x += alpha p ; r -= alpha q ; ||x||^2, ||r||^2
  T* x; T* r; const T* p; const T* q; T alpha;
  __device__ __forceinline__ void operator()(int j, T* d) const {
    const T xn = add_rn(x[j], mul_rn(alpha, p[j]));
    x[j] = xn;
    const T rn = add_rn(r[j], mul_rn(-alpha, q[j]));
    r[j] = rn;
    d[0] += xn * xn; d[1] += rn * rn;
  }
};
template <class T> struct CrK2Epi {                // Ar = A r ; ||Ar||^2, <r, Ar>
  T* Ar; const T* r;
  __device__ __forceinline__ void operator()(int row, T acc, T* d) const { Ar[row] = acc; d[0] += acc * acc; d[1] += __ldg(&r[row]) * acc; }
};
template <class T> struct CrK3Body {               // p = r + beta p ; q = Ar + beta q ; ||q||^2 (the next alpha's denominator)
  T* p; T* q; const T* r; const T* Ar; T beta;
  __device__ __forceinline__ void operator()(int j, T* d) const {
    p[j] = add_rn(r[j], mul_rn(beta, p[j]));
    const T qn = add_rn(Ar[j], mul_rn(beta, q[j]));
    q[j] = qn;
    d[0] += qn * qn;
  }
};
template <class T> void cr_fused_step(Workspace<T>& ws, const Csr<T>& A, T alpha, T* xx, T* rr, T* ArAr, T* rAr) {
  Ctx& c = ws.ctx;
  launch_stream<T, 2>(c, ws.n, CrK1Body<T>{ws.x, ws.r, ws.p, ws.q, alpha}, StoreFin<T, 2>{sib_slots<T>(c)}, 5);
  launch_spmv_epi<T, 2>(c, A, ws.r, CrK2Epi<T>{ws.Ap, ws.r}, StoreFin<T, 2>{sib_slots<T>(c) + 2}, 4);
  T out[4]; sib_read<T, 4>(c, out);
  *xx = out[0]; *rr = out[1]; *ArAr = out[2]; *rAr = out[3];
}
template <class T> T cr_fused_directions(Workspace<T>& ws, T beta) {
  Ctx& c = ws.ctx;
  launch_stream<T, 1>(c, ws.n, CrK3Body<T>{ws.p, ws.q, ws.r, ws.Ap, beta}, StoreFin<T, 1>{sib_slots<T>(c)}, 5);
  T out[1]; sib_read<T, 1>(c, out);
  return out[0];
}

int gmres_fused_max() { return kGmresMaxFused; }

#define INST(T)                                                                                                      \
  template void bicgstab_fused_iteration<T>(Workspace<T>&, const Csr<T>&, const T*, bool, T, T*, T*, T*, T*);         \
  template void minres_fused_lanczos<T>(Workspace<T>&, const Csr<T>&, int, T, T, T, T, T, T, T, T*, T*, T*);          \
  template T minres_fused_update<T>(Workspace<T>&, T*, T, T);                                                        \
  template void gmres_fused_arnoldi<T>(Workspace<T>&, const Csr<T>&, int, T*, T*, const T*);                         \
  template void fused_orth_chain<T>(Workspace<T>&, const Csr<T>&, const T*, T*, const T* const*, int, T*, T*);       \
  template void trunc_fused_direction<T>(Workspace<T>&, T*, int, T* const*, const T*, const T*, T, T);               \
  template T cgs_fused_sigma<T>(Workspace<T>&, const Csr<T>&, const T*);                                             \
  template void cgs_fused_update<T>(Workspace<T>&, const Csr<T>&, const T*, T, T*, T*);                              \
  template void cgs_fused_directions<T>(Workspace<T>&, T);                                                           \
  template T lanczos_fused_delta<T>(Workspace<T>&, const Csr<T>&);                                                   \
  template T lanczos_fused_recur<T>(Workspace<T>&, T, T, bool);                                                      \
  template void lanczos_fused_update<T>(Workspace<T>&, T, T, T, T);                                                  \
  template void cr_fused_step<T>(Workspace<T>&, const Csr<T>&, T, T*, T*, T*, T*);                                   \
  template T cr_fused_directions<T>(Workspace<T>&, T);                                                               \
  template void fused_multi_axpy<T>(Workspace<T>&, T*, int, const T*, T* const*);                                    \
  template void gmres_fused_update_x<T>(Workspace<T>&, T*, int, const T*);
INST(double)
INST(float)
#undef INST

}  // namespace kb
