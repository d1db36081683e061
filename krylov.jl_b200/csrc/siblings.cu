// siblings.cu -- sibling solvers that run on the hot-path kernels unchanged
// (SURVEY.md section 8f-3): cgs!, cg_lanczos!, fom!, fgmres!.  Same rules as
// solvers.cu: the reference's host control flow statement by statement (files
// cited per function), every vector operation a kernel from blas1.cu / spmv.cu
// / fused_phases.cu.  FOM and FGMRES share GMRES's fused Arnoldi step.
#include <cstring>

#include "solver_common.h"

namespace kb {

// ===========================================================================
// cgs!  (src/cgs.jl:125-282)
// ===========================================================================
template <class T>
void cgs_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const T* c_in, const LinOp<T>& M, const LinOp<T>& N,
               const SolveOpts& o) {
  const double start_time = now_seconds();
  Ctx& c = ws.ctx;
  const int n = ws.n;
  const bool history = o.history, ldiv = o.ldiv;
  if (o.verbose > 0) printf("CGS: system of size %d\n", n);
  const bool MisI = M.is_identity(), NisI = N.is_identity();
  allocate_if(!MisI, ws, ws.vw);
  allocate_if(!NisI, ws, ws.yz);
  T *dx = ws.dx, *x = ws.x, *r = ws.r, *u = ws.u, *p = ws.p, *q = ws.q;
  Stats& stats = ws.stats;
  const bool warm_start = ws.warm_start;
  stats.reset();
  T* t = ws.ts; T* s = ws.ts;                                 // cgs.jl:150-155
  T* v = MisI ? t : ws.vw;
  T* w = MisI ? s : ws.vw;
  T* y = NisI ? p : ws.yz;
  T* z = NisI ? u : ws.yz;
  T* r0 = MisI ? r : ws.ts;
  const T* cvec = c_in ? c_in : b;

  if (warm_start) { op_apply(c, A, dx, r0); k_axpby<T>(c, n, T(1), b, T(-1), r0); }
  else k_copy<T>(c, n, r0, b);
  k_fill<T>(c, n, x, T(0));
  if (!MisI) op_apply(c, M, r0, r, ldiv);
  T rNorm = k_nrm2<T>(c, n, r);
  if (history) stats.residuals.push_back(rNorm);
  auto finish_early = [&](bool solved, const char* status) {
    stats.niter = 0; stats.solved = solved; stats.inconsistent = false;
    stats.timer = now_seconds() - start_time; stats.status = status;
    if (warm_start) k_axpy<T>(c, n, T(1), dx, x);
    ws.warm_start = false;
    c.sync();
  };
  if (rNorm == 0) { finish_early(true, "x is a zero-residual solution"); return; }
  T rho = k_dot<T>(c, n, cvec, r);
  if (rho == 0) { finish_early(false, "Breakdown bᴴc = 0"); return; }
  int iter = 0;
  const int itmax = default_itmax(ws, o.itmax);
  const T eps_tol = tol_of<T>(o.atol) + tol_of<T>(o.rtol) * rNorm;
  if (o.verbose > 0) printf("%5s  %7s  %5s\n", "k", "‖rₖ‖", "timer");
  if (kdisplay(iter, o.verbose)) printf("%5d  %7.1e  %.2fs\n", iter, (double)rNorm, now_seconds() - start_time);
  k_copy<T>(c, n, u, r);
  k_copy<T>(c, n, p, r);
  k_fill<T>(c, n, q, T(0));
  bool solved = rNorm <= eps_tol, tired = iter >= itmax, breakdown = false, user_exit = false, overtimed = false;
  std::string status = "unknown";

  // grouped passes (fused_phases.cu): 4 launches and 2 read-backs per iteration instead of 12 and 3
  const bool fusedS = o.fused && A.kind == LinOp<T>::CSR && MisI && NisI;
  while (!(solved || tired || breakdown || user_exit || overtimed)) {
    T alpha, rho_next;
    if (fusedS) {
      const T sigma = cgs_fused_sigma<T>(ws, *A.csr, cvec);
      alpha = rho / sigma;
      T rr;
      cgs_fused_update<T>(ws, *A.csr, cvec, alpha, &rho_next, &rr);
      const T beta = rho_next / rho;
      cgs_fused_directions<T>(ws, beta);
      rho = rho_next;
      iter = iter + 1;
      rNorm = std::sqrt(rr);
    } else {
      if (!NisI) op_apply(c, N, p, y, ldiv);
      op_apply(c, A, y, t);
      if (!MisI) op_apply(c, M, t, v, ldiv);
      const T sigma = k_dot<T>(c, n, cvec, v);
      alpha = rho / sigma;
      k_copy<T>(c, n, q, u);
      k_axpy<T>(c, n, -alpha, v, q);
      k_axpy<T>(c, n, T(1), q, u);
      if (!NisI) op_apply(c, N, u, z, ldiv);
      k_axpy<T>(c, n, alpha, z, x);
      op_apply(c, A, z, s);
      if (!MisI) op_apply(c, M, s, w, ldiv);
      k_axpy<T>(c, n, -alpha, w, r);
      rho_next = k_dot<T>(c, n, cvec, r);
      const T beta = rho_next / rho;
      k_copy<T>(c, n, u, r);
      k_axpy<T>(c, n, beta, q, u);
      k_axpby<T>(c, n, T(1), q, beta, p);
      k_axpby<T>(c, n, T(1), u, beta, p);
      rho = rho_next;
      iter = iter + 1;
      rNorm = k_nrm2<T>(c, n, r);
    }
    if (history) stats.residuals.push_back(rNorm);
    const bool resid_decrease_mach = (rNorm + T(1) <= T(1));
    if (o.callback) { c.sync(); stats.niter = iter; user_exit = o.callback(&ws, o.callback_user) != 0; }
    solved = (rNorm <= eps_tol) || resid_decrease_mach;
    tired = iter >= itmax;
    breakdown = (alpha == 0 || std::isnan(alpha));
    overtimed = (now_seconds() - start_time) > o.timemax;
    agree_exit(ws, o, user_exit, overtimed);      // row-partitioned: same decision on every rank
    if (kdisplay(iter, o.verbose)) printf("%5d  %7.1e  %.2fs\n", iter, (double)rNorm, now_seconds() - start_time);
  }
  if (o.verbose > 0) printf("\n");
  if (tired) status = "maximum number of iterations exceeded";
  if (breakdown) status = "breakdown αₖ == 0";
  if (solved) status = "solution good enough given atol and rtol";
  if (user_exit) status = "user-requested exit";
  if (overtimed) status = "time limit exceeded";
  if (warm_start) k_axpy<T>(c, n, T(1), dx, x);
  ws.warm_start = false;
  c.sync();
  stats.niter = iter; stats.solved = solved; stats.inconsistent = false;
  stats.timer = now_seconds() - start_time;
  stats.status = status;
}

// ===========================================================================
// cg_lanczos!  (src/cg_lanczos.jl:110-264)
// ===========================================================================
template <class T>
void cg_lanczos_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const SolveOpts& o) {
  const double start_time = now_seconds();
  Ctx& c = ws.ctx;
  const int n = ws.n;
  const bool history = o.history, ldiv = o.ldiv, check_curvature = o.check_curvature;
  if (o.verbose > 0) printf("CG-LANCZOS: system of %d equations in %d variables\n", n, n);
  const bool MisI = M.is_identity();
  allocate_if(!MisI, ws, ws.vv);
  T *dx = ws.dx, *x = ws.x, *Mv = ws.Mv, *Mv_prev = ws.Mv_prev, *p = ws.p, *Mv_next = ws.Mv_next;
  Stats& stats = ws.stats;
  const bool warm_start = ws.warm_start;
  stats.reset();
  stats.Anorm = NAN;
  T* v = MisI ? Mv : ws.vv;                                   // cg_lanczos.jl:138
  // knorm_elliptic (src/krylov_utils.jl:319): ||v|| when v === Mv, else sqrt(<v, Mv>)
  auto norm_elliptic = [&]() { return MisI ? k_nrm2<T>(c, n, v) : (T)std::sqrt(k_dot<T>(c, n, v, Mv)); };

  k_fill<T>(c, n, x, T(0));
  if (warm_start) { op_apply(c, A, dx, Mv); k_axpby<T>(c, n, T(1), b, T(-1), Mv); }
  else k_copy<T>(c, n, Mv, b);
  if (!MisI) op_apply(c, M, Mv, v, ldiv);
  T beta = norm_elliptic();
  T sigma = beta;
  T rNorm = sigma;
  if (history) stats.residuals.push_back(rNorm);
  if (beta == 0) {
    stats.niter = 0; stats.solved = true; stats.Anorm = 0; stats.indefinite = false;
    stats.timer = now_seconds() - start_time;
    stats.status = "x is a zero-residual solution";
    if (warm_start) k_axpy<T>(c, n, T(1), dx, x);
    ws.warm_start = false;
    c.sync();
    return;
  }
  k_copy<T>(c, n, p, v);
  k_scal<T>(c, n, T(1) / beta, v);                            // kdiv!(n, v, β)
  if (!MisI) k_scal<T>(c, n, T(1) / beta, Mv);
  k_copy<T>(c, n, Mv_prev, Mv);
  int iter = 0;
  const int itmax = default_itmax(ws, o.itmax);
  T omega = 0, gamma = 1, Anorm2 = 0, beta_prev = 0;
  const T eps_tol = tol_of<T>(o.atol) + tol_of<T>(o.rtol) * rNorm;
  if (o.verbose > 0) printf("%5s  %7s  %5s\n", "k", "‖rₖ‖", "timer");
  if (kdisplay(iter, o.verbose)) printf("%5d  %7.1e  %.2fs\n", iter, (double)rNorm, now_seconds() - start_time);
  bool indefinite = false, solved = rNorm <= eps_tol, tired = iter >= itmax, user_exit = false, overtimed = false;
  std::string status = "unknown";

  while (!(solved || tired || (check_curvature && indefinite) || user_exit || overtimed)) {
    // grouped passes (fused_phases.cu; M = I so v === Mv): 3 launches per iteration instead of 8
    const bool fusedL = o.fused && A.kind == LinOp<T>::CSR && MisI;
    const T delta = fusedL ? lanczos_fused_delta<T>(ws, *A.csr) : (op_apply(c, A, v, Mv_next), k_dot<T>(c, n, v, Mv_next));
    gamma = T(1) / (delta - omega / gamma);
    indefinite = indefinite || (gamma <= 0);
    if (check_curvature && indefinite) continue;
    if (fusedL) {
      beta = lanczos_fused_recur<T>(ws, delta, beta, iter > 0);
    } else {
      k_axpy<T>(c, n, -delta, Mv, Mv_next);
      if (iter > 0) {
        k_axpy<T>(c, n, -beta, Mv_prev, Mv_next);
        k_copy<T>(c, n, Mv_prev, Mv);
    }
    k_copy<T>(c, n, Mv, Mv_next);
    if (!MisI) op_apply(c, M, Mv, v, ldiv);
    beta = norm_elliptic();
    k_scal<T>(c, n, T(1) / beta, v);
    if (!MisI) k_scal<T>(c, n, T(1) / beta, Mv);
    }
    Anorm2 += beta_prev * beta_prev + beta * beta + delta * delta;
    beta_prev = beta;
    if (!fusedL) k_axpy<T>(c, n, gamma, p, x);
    omega = beta * gamma;
    sigma = -omega * sigma;
    omega = omega * omega;
    if (fusedL) lanczos_fused_update<T>(ws, beta, gamma, sigma, omega);
    else k_axpby<T>(c, n, sigma, v, omega, p);
    rNorm = std::fabs(sigma);
    if (history) stats.residuals.push_back(rNorm);
    iter = iter + 1;
    if (kdisplay(iter, o.verbose)) printf("%5d  %7.1e  %.2fs\n", iter, (double)rNorm, now_seconds() - start_time);
    const bool resid_decrease_mach = (rNorm + T(1) <= T(1));
    if (o.callback) { c.sync(); stats.niter = iter; user_exit = o.callback(&ws, o.callback_user) != 0; }
    solved = (rNorm <= eps_tol) || resid_decrease_mach;
    tired = iter >= itmax;
    overtimed = (now_seconds() - start_time) > o.timemax;
    agree_exit(ws, o, user_exit, overtimed);      // row-partitioned: same decision on every rank
  }
  if (o.verbose > 0) printf("\n");
  if (tired) status = "maximum number of iterations exceeded";
  if (check_curvature && indefinite) status = "negative curvature";
  if (solved) status = "solution good enough given atol and rtol";
  if (user_exit) status = "user-requested exit";
  if (overtimed) status = "time limit exceeded";
  if (warm_start) k_axpy<T>(c, n, T(1), dx, x);
  ws.warm_start = false;
  c.sync();
  stats.niter = iter; stats.solved = solved; stats.Anorm = std::sqrt(Anorm2); stats.indefinite = indefinite;
  stats.timer = now_seconds() - start_time;
  stats.status = status;
}

// ---------------------------------------------------------------------------
// fom! and fgmres! share everything but the small factorization kept on the
// host (LU vs Givens QR) and where the right preconditioner lives (FGMRES
// stores Z[k] = N V[k]); one driver, the differences are marked.
//   fom!     src/fom.jl:121-368
//   fgmres!  src/fgmres.jl:128-388
// ---------------------------------------------------------------------------
template <class T, bool FLEX>
static void arnoldi_family_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const LinOp<T>& N,
                                 const SolveOpts& o) {
  const double start_time = now_seconds();
  Ctx& cx = ws.ctx;
  const int n = ws.n;
  const bool history = o.history, ldiv = o.ldiv, restart = o.restart, reorth = o.reorthogonalization;
  if (o.verbose > 0) printf("%s: system of size %d\n", FLEX ? "FGMRES" : "FOM", n);
  const bool MisI = M.is_identity(), NisI = N.is_identity();
  allocate_if(!MisI, ws, ws.q);
  if (!FLEX) allocate_if(!NisI, ws, ws.pp);
  allocate_if(restart, ws, ws.dx);
  T *dx = ws.dx, *x = ws.x, *w = ws.w;
  std::vector<T*>& V = ws.V;
  std::vector<T*>& Z = ws.Z;
  // FGMRES: c, s (sgiv), z (zg), R.   FOM: l (sgiv), z (zg), U (R); c is unused.
  std::vector<T>&c = ws.c, &s = ws.sgiv, &z = ws.zg, &R = ws.R;
  std::vector<T>& l = ws.sgiv;
  Stats& stats = ws.stats;
  const bool warm_start = ws.warm_start;
  stats.reset();
  T* q = MisI ? w : ws.q;
  T* r0 = MisI ? w : ws.q;
  T* xr = restart ? dx : x;

  k_fill<T>(cx, n, x, T(0));
  if (warm_start) {
    op_apply(cx, A, dx, w);
    k_axpby<T>(cx, n, T(1), b, T(-1), w);
    if (restart) k_axpy<T>(cx, n, T(1), dx, x);
  } else {
    k_copy<T>(cx, n, w, b);
  }
  if (!MisI) op_apply(cx, M, w, r0, ldiv);
  T beta = k_nrm2<T>(cx, n, r0);
  T rNorm = beta;
  if (history) stats.residuals.push_back(beta);
  const T eps_tol = tol_of<T>(o.atol) + tol_of<T>(o.rtol) * rNorm;
  if (beta == 0) {
    stats.niter = 0; stats.solved = true; stats.inconsistent = false;
    stats.timer = now_seconds() - start_time;
    stats.status = "x is a zero-residual solution";
    if (warm_start) k_axpy<T>(cx, n, T(1), dx, x);
    ws.warm_start = false;
    cx.sync();
    return;
  }
  const int mem = (int)s.size();                              // length(c) / length(l)
  int npass = 0, iter = 0, inner_iter = 0;
  const int itmax = default_itmax(ws, o.itmax);
  int inner_itmax = itmax;
  if (o.verbose > 0) printf("%5s  %5s  %7s  %7s  %5s\n", "pass", "k", "‖rₖ‖", "hₖ₊₁.ₖ", "timer");
  if (kdisplay(iter, o.verbose)) printf("%5d  %5d  %7.1e  %7s  %.2fs\n", npass, iter, (double)rNorm, "✗ ✗ ✗ ✗", now_seconds() - start_time);
  const T btol = std::pow(eps_of<T>(), T(0.75));
  // the fused Arnoldi step folds a left diagonal M into the SpMV epilogue; FOM needs N = I (FGMRES materialises Z[k])
  const bool fusedA = o.fused && A.kind == LinOp<T>::CSR && !reorth && (FLEX || NisI) &&
                      (MisI || (M.kind == LinOp<T>::DIAG && !ldiv));
  ws.mdiag_fused = (fusedA && !MisI) ? M.diag : nullptr;
  bool breakdown = false, inconsistent = false, solved = rNorm <= eps_tol, tired = iter >= itmax;
  bool inner_tired = inner_iter >= inner_itmax, user_exit = false, overtimed = false;
  std::string status = "unknown";

  while (!(solved || tired || breakdown || user_exit || overtimed)) {
    int nr = 0;
    // The reference zero-fills V (and Z) every cycle.  FGMRES reads only entries it wrote first, so (as in
    // gmres_solve) the fill is kept only where callbacks could see unused columns.  FOM does read a zero V[k+1]
    // after a user exit or timeout (its inner loop does not test them, fom.jl:237), so FOM always fills.
    if (!restart || !FLEX) for (int i = 0; i < mem; i++) { k_fill<T>(cx, n, V[i], T(0)); if (FLEX) k_fill<T>(cx, n, Z[i], T(0)); }
    std::fill(s.begin(), s.end(), T(0));
    if (FLEX) std::fill(c.begin(), c.end(), T(0));
    std::fill(R.begin(), R.end(), T(0));
    std::fill(z.begin(), z.end(), T(0));
    if (restart) {
      k_fill<T>(cx, n, xr, T(0));
      if (npass >= 1) {
        op_apply(cx, A, x, w);
        k_axpby<T>(cx, n, T(1), b, T(-1), w);
        if (!MisI) op_apply(cx, M, w, r0, ldiv);
      }
    }
    beta = k_nrm2<T>(cx, n, r0);
    z[0] = beta;
    k_divcopy<T>(cx, n, V[0], r0, rNorm);
    npass = npass + 1;
    ws.inner_iter = 0;
    inner_tired = false;

    // fom.jl:237 tests only solved/inner_tired/breakdown; fgmres.jl:243 also the user exit and the timer
    while (!(solved || inner_tired || breakdown || (FLEX && (user_exit || overtimed)))) {
      ws.inner_iter = ws.inner_iter + 1;
      inner_iter = ws.inner_iter;
      if (!restart && (inner_iter > mem)) {
        const double t0 = now_seconds();
        for (int i = 0; i < inner_iter; i++) R.push_back(T(0));
        s.push_back(T(0));                                    // FGMRES s / FOM l
        if (FLEX) { c.push_back(T(0)); Z.push_back(dev_alloc<T>((size_t)n)); }
        else z.push_back(T(0));                               // fom.jl:249 grows z here, fgmres.jl:329 with V
        stats.allocation_timer += now_seconds() - t0;
      }
      T* vk = V[inner_iter - 1];
      T* p;
      if (FLEX) {                                             // z_k <- N_k v_k, unconditional (fgmres.jl:262)
        p = Z[inner_iter - 1];
        if (NisI) k_copy<T>(cx, n, p, vk); else op_apply(cx, N, vk, p, ldiv);
      } else {
        p = NisI ? vk : ws.pp;
        if (!NisI) op_apply(cx, N, vk, p, ldiv);
      }
      T Hbis;
      if (fusedA && inner_iter <= gmres_fused_max()) {
        gmres_fused_arnoldi<T>(ws, *A.csr, inner_iter, &R[nr], &Hbis, p);
      } else {
        op_apply(cx, A, p, w);
        if (!MisI) op_apply(cx, M, w, q, ldiv);
        for (int i = 0; i < inner_iter; i++) {
          R[nr + i] = k_dot<T>(cx, n, V[i], q);
          k_axpy<T>(cx, n, -R[nr + i], V[i], q);
        }
        if (reorth) {
          for (int i = 0; i < inner_iter; i++) {
            const T Htmp = k_dot<T>(cx, n, V[i], q);
            R[nr + i] += Htmp;
            k_axpy<T>(cx, n, -Htmp, V[i], q);
          }
        }
        Hbis = k_nrm2<T>(cx, n, q);
      }
      T zeta_next = 0;
      if (FLEX) {                                             // Givens QR of H, fgmres.jl:285-303
        for (int i = 0; i < inner_iter - 1; i++) {
          const T Rtmp = c[i] * R[nr + i] + s[i] * R[nr + i + 1];
          R[nr + i + 1] = s[i] * R[nr + i] - c[i] * R[nr + i + 1];
          R[nr + i] = Rtmp;
        }
        sym_givens<T>(R[nr + inner_iter - 1], Hbis, &c[inner_iter - 1], &s[inner_iter - 1], &R[nr + inner_iter - 1]);
        zeta_next = s[inner_iter - 1] * z[inner_iter - 1];
        z[inner_iter - 1] = c[inner_iter - 1] * z[inner_iter - 1];
        rNorm = std::fabs(zeta_next);
      } else {                                                // LU of H without pivoting, fom.jl:274-288
        if (inner_iter >= 2) {
          for (int i = 2; i <= inner_iter; i++) R[nr + i - 1] = R[nr + i - 1] - l[i - 2] * R[nr + i - 2];
          z[inner_iter - 1] = -l[inner_iter - 2] * z[inner_iter - 2];
        }
        l[inner_iter - 1] = Hbis / R[nr + inner_iter - 1];
        rNorm = Hbis * std::fabs(z[inner_iter - 1] / R[nr + inner_iter - 1]);
      }
      if (history) stats.residuals.push_back(rNorm);
      nr = nr + inner_iter;
      const bool resid_decrease_mach = (rNorm + T(1) <= T(1));
      if (o.callback) { cx.sync(); stats.niter = iter + inner_iter; user_exit = o.callback(&ws, o.callback_user) != 0; }
      const bool resid_decrease_lim = rNorm <= eps_tol;
      breakdown = Hbis <= btol;
      solved = resid_decrease_lim || resid_decrease_mach;
      inner_tired = restart ? inner_iter >= std::min(mem, inner_itmax) : inner_iter >= inner_itmax;
      overtimed = (now_seconds() - start_time) > o.timemax;
      agree_exit(ws, o, user_exit, overtimed);      // row-partitioned: same decision on every rank
      if (kdisplay(iter + inner_iter, o.verbose))
        printf("%5d  %5d  %7.1e  %7.1e  %.2fs\n", npass, iter + inner_iter, (double)rNorm, (double)Hbis, now_seconds() - start_time);
      if (!(solved || inner_tired || breakdown || user_exit || overtimed)) {
        if (!restart && (inner_iter >= mem)) {
          const double t0 = now_seconds();
          V.push_back(dev_alloc<T>((size_t)n));
          if (FLEX) z.push_back(T(0));
          stats.allocation_timer += now_seconds() - t0;
        }
        k_divcopy<T>(cx, n, V[inner_iter], q, Hbis);
        if (FLEX) z[inner_iter] = zeta_next;
      }
    }
    std::vector<T>& y = z;                                    // back substitution, fom.jl:322-331 / fgmres.jl:335-350
    for (int i = inner_iter; i >= 1; i--) {
      int pos = nr + i - inner_iter;
      for (int j = inner_iter; j >= i + 1; j--) {
        y[i - 1] = y[i - 1] - R[pos - 1] * y[j - 1];
        pos = pos - j + 1;
      }
      if (FLEX && std::fabs(R[pos - 1]) <= btol) { y[i - 1] = T(0); inconsistent = true; }
      else y[i - 1] = y[i - 1] / R[pos - 1];
    }
    T* const* basis = FLEX ? Z.data() : V.data();             // x_k = Z_k y_k (FGMRES) or N V_k y_k (FOM)
    if (fusedA) fused_multi_axpy<T>(ws, xr, inner_iter, y.data(), basis);
    else for (int i = 0; i < inner_iter; i++) k_axpy<T>(cx, n, y[i], basis[i], xr);
    if (!FLEX && !NisI) { k_copy<T>(cx, n, ws.pp, xr); op_apply(cx, N, ws.pp, xr, ldiv); }
    if (restart) k_axpy<T>(cx, n, T(1), xr, x);
    inner_itmax = inner_itmax - inner_iter;
    iter = iter + inner_iter;
    tired = iter >= itmax;
    overtimed = (now_seconds() - start_time) > o.timemax;
    agree_exit(ws, o, user_exit, overtimed);      // row-partitioned: same decision on every rank
  }
  if (o.verbose > 0) printf("\n");
  if (tired) status = "maximum number of iterations exceeded";
  if (!FLEX && breakdown) status = "inconsistent linear system";
  if (solved) status = "solution good enough given atol and rtol";
  if (FLEX && inconsistent) status = "found approximate least-squares solution";
  if (user_exit) status = "user-requested exit";
  if (overtimed) status = "time limit exceeded";
  if (warm_start && !restart) k_axpy<T>(cx, n, T(1), dx, x);
  ws.warm_start = false;
  cx.sync();
  stats.niter = iter; stats.solved = solved;
  stats.inconsistent = FLEX ? inconsistent : (!solved && breakdown);
  stats.timer = now_seconds() - start_time;
  stats.status = status;
}

template <class T>
void fom_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const LinOp<T>& N, const SolveOpts& o) {
  arnoldi_family_solve<T, false>(ws, A, b, M, N, o);
}
template <class T>
void fgmres_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const LinOp<T>& N, const SolveOpts& o) {
  arnoldi_family_solve<T, true>(ws, A, b, M, N, o);
}

// ---------------------------------------------------------------------------
// dqgmres! (src/dqgmres.jl:121-335) and diom! (src/diom.jl:121-332): the truncated
// (incomplete orthogonalization) variants; circular stacks V, P of `memory` vectors.
// One driver: QR by Givens rotations (DQGMRES) or LU without pivoting (DIOM) of the
// band Hessenberg matrix, kept on the host.  Indices are 1-based like the reference.
//   workspace fields: t (ws.t), z (ws.yz), w (ws.vw), V, P (ws.Z), c, s / L (ws.sgiv), H (ws.R)
// ---------------------------------------------------------------------------
template <class T, bool QR>
static void truncated_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const LinOp<T>& N, const SolveOpts& o) {
  const double start_time = now_seconds();
  Ctx& cx = ws.ctx;
  const int n = ws.n;
  const bool history = o.history, ldiv = o.ldiv, reorth = o.reorthogonalization;
  if (o.verbose > 0) printf("%s: system of size %d\n", QR ? "DQGMRES" : "DIOM", n);
  const bool MisI = M.is_identity(), NisI = N.is_identity();
  allocate_if(!MisI, ws, ws.vw);
  allocate_if(!NisI, ws, ws.yz);
  T *dx = ws.dx, *x = ws.x, *t = ws.t;
  std::vector<T*>&P = ws.Z, &V = ws.V;
  std::vector<T>&c = ws.c, &s = ws.sgiv, &H = ws.R;
  std::vector<T>& L = ws.sgiv;
  Stats& stats = ws.stats;
  const bool warm_start = ws.warm_start;
  stats.reset();
  T* w = MisI ? t : ws.vw;
  T* r0 = MisI ? t : ws.vw;

  k_fill<T>(cx, n, x, T(0));
  if (warm_start) { op_apply(cx, A, dx, t); k_axpby<T>(cx, n, T(1), b, T(-1), t); }
  else k_copy<T>(cx, n, t, b);
  if (!MisI) op_apply(cx, M, t, r0, ldiv);
  T rNorm = k_nrm2<T>(cx, n, r0);
  if (history) stats.residuals.push_back(rNorm);
  if (rNorm == 0) {
    stats.niter = 0; stats.solved = true; stats.inconsistent = false;
    stats.timer = now_seconds() - start_time;
    stats.status = "x is a zero-residual solution";
    if (warm_start) k_axpy<T>(cx, n, T(1), dx, x);
    ws.warm_start = false;
    cx.sync();
    return;
  }
  int iter = 0;
  const int itmax = default_itmax(ws, o.itmax);
  const T eps_tol = tol_of<T>(o.atol) + tol_of<T>(o.rtol) * rNorm;
  if (o.verbose > 0) printf("%5s  %7s  %5s\n", "k", "‖rₖ‖", "timer");
  if (kdisplay(iter, o.verbose)) printf("%5d  %7.1e  %.2fs\n", iter, (double)rNorm, now_seconds() - start_time);
  const int mem = (int)V.size();
  for (int i = 0; i < mem; i++) k_fill<T>(cx, n, V[i], T(0));
  for (size_t i = 0; i < P.size(); i++) k_fill<T>(cx, n, P[i], T(0));
  std::fill(H.begin(), H.end(), T(0));
  std::fill(s.begin(), s.end(), T(0));          // DQGMRES sines / DIOM pivots L
  if (QR) std::fill(c.begin(), c.end(), T(0));
  T gamma_k = rNorm;                            // DQGMRES: last component of g_k;  DIOM: xi
  k_divcopy<T>(cx, n, V[0], r0, rNorm);
  bool solved = rNorm <= eps_tol, tired = iter >= itmax, user_exit = false, overtimed = false;
  std::string status = "unknown";

  // grouped passes (fused_phases.cu): window + 3 launches and one read-back per iteration instead of 2 window + 6
  constexpr int kMaxWindow = 120;
  const bool fusedT = o.fused && A.kind == LinOp<T>::CSR && MisI && NisI && !reorth && mem <= kMaxWindow && mem <= gmres_fused_max();
  ws.mdiag_fused = nullptr;
  T* dvec[kMaxWindow]; T dcoef[kMaxWindow];
  while (!(solved || tired || user_exit || overtimed)) {
    iter = iter + 1;
    int ndir = 0;
    const int pos = (iter - 1) % mem + 1, next_pos = iter % mem + 1;
    T* z = NisI ? V[pos - 1] : ws.yz;
    const int lo = std::max(1, iter - mem + 1);
    T Haux;
    if (fusedT) {
      // SpMV + incomplete orthogonalization as ONE chain of passes (fused_phases.cu: fused_orth_chain): the dot
      // product with the next basis vector rides in the pass that subtracts the current one; one read-back
      const T* vecs[kMaxWindow]; T hbuf[kMaxWindow];
      const int cnt = iter - lo + 1;
      for (int i = lo; i <= iter; i++) vecs[i - lo] = V[(i - 1) % mem];
      fused_orth_chain<T>(ws, *A.csr, z, w, vecs, cnt, hbuf, &Haux);
      for (int i = lo; i <= iter; i++) H[iter - i] = hbuf[i - lo];
    } else {
      if (!NisI) op_apply(cx, N, V[pos - 1], z, ldiv);
      op_apply(cx, A, z, t);
      if (!MisI) op_apply(cx, M, t, w, ldiv);
      for (int i = lo; i <= iter; i++) {          // incomplete orthogonalization
        const int ipos = (i - 1) % mem + 1, diag = iter - i + 1;
        H[diag - 1] = k_dot<T>(cx, n, w, V[ipos - 1]);
        k_axpy<T>(cx, n, -H[diag - 1], V[ipos - 1], w);
    }
    if (reorth) {
      for (int i = lo; i <= iter; i++) {
        const int ipos = (i - 1) % mem + 1, diag = iter - i + 1;
        const T Htmp = k_dot<T>(cx, n, w, V[ipos - 1]);
        H[diag - 1] += Htmp;
        k_axpy<T>(cx, n, -Htmp, V[ipos - 1], w);
      }
    }
    Haux = k_nrm2<T>(cx, n, w);
    }
    if (Haux != 0) k_divcopy<T>(cx, n, V[next_pos - 1], w, Haux);
    int ppos;                                   // position of p_k in the circular stack P
    T step;                                     // x += step * p_k
    if (QR) {                                   // dqgmres.jl:268-289
      if (iter >= mem + 2) H[mem] = T(0);
      const int lo2 = std::max(1, iter - mem);
      for (int i = lo2; i <= iter - 1; i++) {
        const int irot = (i - 1) % mem + 1, diag = iter - i, next_diag = diag + 1;
        const T Htmp = c[irot - 1] * H[next_diag - 1] + s[irot - 1] * H[diag - 1];
        H[diag - 1] = s[irot - 1] * H[next_diag - 1] - c[irot - 1] * H[diag - 1];
        H[next_diag - 1] = Htmp;
      }
      sym_givens<T>(H[0], Haux, &c[pos - 1], &s[pos - 1], &H[0]);
      const T gamma_next = s[pos - 1] * gamma_k;
      gamma_k = c[pos - 1] * gamma_k;
      ppos = pos;
      for (int i = lo2; i <= iter - 1; i++) {
        const int ipos = (i - 1) % mem + 1, diag = iter - i + 1;
        if (fusedT) { dvec[ndir] = P[ipos - 1]; dcoef[ndir++] = -H[diag - 1]; }
        else if (ipos == ppos) k_scal<T>(cx, n, -H[diag - 1], P[ppos - 1]);
        else k_axpy<T>(cx, n, -H[diag - 1], P[ipos - 1], P[ppos - 1]);
      }
      step = gamma_k;
      rNorm = std::fabs(gamma_next);
      gamma_k = gamma_next;
    } else {                                    // diom.jl:262-289
      if (iter >= 2) {
        for (int i = std::max(2, iter - mem + 2); i <= iter; i++) {
          const int lpos = (i - 1) % (mem - 1) + 1, diag = iter - i + 1, next_diag = diag + 1;
          H[diag - 1] = H[diag - 1] - L[lpos - 1] * H[next_diag - 1];
          if (i == iter) gamma_k = -L[lpos - 1] * gamma_k;
        }
      }
      const int next_lpos = iter % (mem - 1) + 1;
      L[next_lpos - 1] = Haux / H[0];
      ppos = (iter - 1) % (mem - 1) + 1;
      for (int i = lo; i <= iter - 1; i++) {
        const int ipos = (i - 1) % (mem - 1) + 1, diag = iter - i + 1;
        if (fusedT) { dvec[ndir] = P[ipos - 1]; dcoef[ndir++] = -H[diag - 1]; }
        else if (ipos == ppos) k_scal<T>(cx, n, -H[diag - 1], P[ppos - 1]);
        else k_axpy<T>(cx, n, -H[diag - 1], P[ipos - 1], P[ppos - 1]);
      }
      step = gamma_k;
      rNorm = Haux * std::fabs(gamma_k / H[0]);
    }
    if (fusedT) {
      // the whole direction update + x update in one pass per 8 stack vectors (trunc_fused_direction)
      trunc_fused_direction<T>(ws, P[ppos - 1], ndir, dvec, dcoef, z, H[0], step);
    } else {
      k_axpy<T>(cx, n, T(1), z, P[ppos - 1]);
      k_scal<T>(cx, n, T(1) / H[0], P[ppos - 1]);   // kdiv!(n, P[pos], H[1])
      k_axpy<T>(cx, n, step, P[ppos - 1], x);
    }
    if (history) stats.residuals.push_back(rNorm);
    const bool resid_decrease_mach = (rNorm + T(1) <= T(1));
    if (o.callback) { cx.sync(); stats.niter = iter; user_exit = o.callback(&ws, o.callback_user) != 0; }
    solved = (rNorm <= eps_tol) || resid_decrease_mach;
    tired = iter >= itmax;
    overtimed = (now_seconds() - start_time) > o.timemax;
    agree_exit(ws, o, user_exit, overtimed);      // row-partitioned: same decision on every rank
    if (kdisplay(iter, o.verbose)) printf("%5d  %7.1e  %.2fs\n", iter, (double)rNorm, now_seconds() - start_time);
  }
  if (o.verbose > 0) printf("\n");
  if (QR) {                                     // dqgmres.jl:319-322 assigns in this order (tired overrides solved)
    if (solved) status = "solution good enough given atol and rtol";
    if (tired) status = "maximum number of iterations exceeded";
  } else {
    if (tired) status = "maximum number of iterations exceeded";
    if (solved) status = "solution good enough given atol and rtol";
  }
  if (user_exit) status = "user-requested exit";
  if (overtimed) status = "time limit exceeded";
  if (warm_start) k_axpy<T>(cx, n, T(1), dx, x);
  ws.warm_start = false;
  cx.sync();
  stats.niter = iter; stats.solved = solved; stats.inconsistent = false;
  stats.timer = now_seconds() - start_time;
  stats.status = status;
}

template <class T>
void dqgmres_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const LinOp<T>& N, const SolveOpts& o) {
  truncated_solve<T, true>(ws, A, b, M, N, o);
}
template <class T>
void diom_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const LinOp<T>& N, const SolveOpts& o) {
  truncated_solve<T, false>(ws, A, b, M, N, o);
}

// ===========================================================================
// cr!  (src/cr.jl:128-478), trust region and linesearch included
//   workspace fields: r, p, q, Ar (ws.Ap), Mq (ws.z), npc_dir
// ===========================================================================
// to_boundary(n, x, d, z, radius; flip = false, xNorm2, dNorm2) with M = I (src/krylov_utils.jl:375-402)
template <class T>
static void cr_to_boundary(Ctx& c, int n, const T* x, const T* d, T radius, T xNorm2, T dNorm2, T* lo, T* hi) {
  if (!(radius > 0)) throw std::runtime_error("radius must be positive");
  const T rxd = k_dot<T>(c, n, x, d);
  if (dNorm2 == T(0)) dNorm2 = k_dot<T>(c, n, d, d);
  if (xNorm2 == T(0)) xNorm2 = k_dot<T>(c, n, x, x);
  if (dNorm2 == T(0)) throw std::runtime_error("zero direction");
  const T radius2 = radius * radius;
  if (!(xNorm2 <= radius2)) throw std::runtime_error("outside of the trust region");
  T s1, s2;
  if (roots_quadratic<T>(dNorm2, 2 * rxd, xNorm2 - radius2, 1, &s1, &s2)) throw std::runtime_error("negative discriminant");
  *hi = std::max(s1, s2); *lo = std::min(s1, s2);
}

template <class T>
void cr_solve(Workspace<T>& ws, const LinOp<T>& A, const T* b, const LinOp<T>& M, const SolveOpts& o) {
  const double start_time = now_seconds();
  Ctx& c = ws.ctx;
  const int n = ws.n;
  const bool history = o.history, ldiv = o.ldiv, linesearch = o.linesearch;
  const T radius = (T)o.radius;
  const T gam = o.cr_gamma < 0 ? std::sqrt(eps_of<T>()) : (T)o.cr_gamma;
  if (linesearch && radius > 0) throw std::runtime_error("'linesearch' set to 'true' but radius > 0");
  if (o.verbose > 0) printf("CR: system of %d equations in %d variables\n", n, n);
  if (ws.warm_start && linesearch) throw std::runtime_error("warm_start and linesearch cannot be used together");
  const bool MisI = M.is_identity();
  allocate_if(!MisI, ws, ws.z);
  allocate_if(linesearch || radius > 0, ws, ws.npc_dir);
  T *dx = ws.dx, *x = ws.x, *r = ws.r, *p = ws.p, *q = ws.q, *Ar = ws.Ap;
  Stats& stats = ws.stats;
  const bool warm_start = ws.warm_start;
  stats.reset();
  T* Mq = MisI ? q : ws.z;
  T* npc_dir = ws.npc_dir;

  k_fill<T>(c, n, x, T(0));
  if (warm_start) { op_apply(c, A, dx, p); k_axpby<T>(c, n, T(1), b, T(-1), p); }
  else k_copy<T>(c, n, p, b);
  if (MisI) k_copy<T>(c, n, r, p); else op_apply(c, M, p, r, ldiv);
  T rNorm = std::sqrt(k_dot<T>(c, n, r, p));                  // knorm_elliptic(n, r, p)
  if (history) stats.residuals.push_back(rNorm);
  if (rNorm == 0) {
    stats.niter = 0; stats.solved = true; stats.inconsistent = false;
    stats.timer = now_seconds() - start_time;
    stats.status = "x is a zero-residual solution";
    if (history) stats.Aresiduals.push_back(0);
    if (warm_start) k_axpy<T>(c, n, T(1), dx, x);
    ws.warm_start = false;
    c.sync();
    return;
  }
  op_apply(c, A, r, Ar);
  T rho = k_dot<T>(c, n, r, Ar);
  if (rho == 0) {
    stats.niter = 0; stats.solved = true; stats.inconsistent = false;
    stats.timer = now_seconds() - start_time;
    stats.status = "b is a zero-curvature direction";
    if (history) stats.Aresiduals.push_back(0);
    ws.warm_start = false;
    if (linesearch || radius > 0) {
      k_copy<T>(c, n, x, p);
      k_copy<T>(c, n, npc_dir, p);
      stats.npcCount = 1; stats.indefinite = true;
    }
    c.sync();
    return;
  }
  k_copy<T>(c, n, p, r);
  k_copy<T>(c, n, q, Ar);
  T mquad = 0;                                                // quadratic model (verbose only)
  int iter = 0;
  const int itmax = default_itmax(ws, o.itmax);
  T rNorm2 = rNorm * rNorm, pNorm = rNorm, pNorm2 = rNorm2, pr = rNorm2, abspr = pr, pAp = rho, abspAp = std::fabs(pAp);
  T xNorm = 0;
  T ArNorm = k_nrm2<T>(c, n, Ar);
  if (history) stats.Aresiduals.push_back(ArNorm);
  const T eps_tol = tol_of<T>(o.atol) + tol_of<T>(o.rtol) * rNorm;
  if (o.verbose > 0) printf("%5s  %8s  %8s  %8s  %5s\n", "k", "‖x‖", "‖r‖", "quad", "timer");
  if (kdisplay(iter, o.verbose)) printf("%5d  %8.1e  %8.1e  %8.1e  %.2fs\n", iter, (double)xNorm, (double)rNorm, (double)mquad, now_seconds() - start_time);
  bool descent = pr > 0, solved = rNorm <= eps_tol, tired = iter >= itmax, on_boundary = false, npcurv = false;
  bool user_exit = false, overtimed = false;
  // grouped passes (no trust region, no linesearch, M = I): 3 launches and 2 read-backs per iteration instead of 9 and 5
  const bool fusedC = o.fused && A.kind == LinOp<T>::CSR && MisI && radius == 0 && !linesearch;
  T qq = 0;
  bool have_qq = false;
  std::string status = "unknown";
  const T sqeps = std::sqrt(eps_of<T>());

  while (!(solved || tired || user_exit || overtimed)) {
    T alpha = 0;
    if (linesearch) {
      const bool p_curv = pAp <= gam * pNorm * pNorm, r_curv = rho <= gam * rNorm * rNorm;
      if (p_curv || r_curv) {                                 // cr.jl:233-262
        npcurv = true;
        if (o.verbose > 0) printf("nonpositive curvature detected: pᴴAp = %8.1e and rᴴAr = %8.1e\n", (double)pAp, (double)rho);
        stats.solved = true; stats.niter = iter; stats.inconsistent = false;
        stats.timer = now_seconds() - start_time;
        stats.status = "nonpositive curvature";
        ws.warm_start = false;
        stats.indefinite = true;
        if (iter == 0) {
          k_copy<T>(c, n, npc_dir, p);
          k_copy<T>(c, n, x, p);
          stats.npcCount = 1;
        } else {
          if (r_curv) { k_copy<T>(c, n, npc_dir, r); stats.npcCount += 1; }
          if (p_curv) { stats.npcCount += 1; if (!r_curv) k_copy<T>(c, n, npc_dir, p); }
        }
        c.sync();
        return;
      }
    } else if (pAp <= 0 && radius == 0) {
      throw std::runtime_error("Indefinite system and no trust region");
    }
    if (!MisI) op_apply(c, M, q, Mq, ldiv);
    if (radius > 0) {                                         // cr.jl:268-373
      const T xNorm2 = xNorm * xNorm;
      T t1, t2, tr, tlo;
      cr_to_boundary<T>(c, n, x, p, radius, xNorm2, pNorm2, &t2, &t1);
      cr_to_boundary<T>(c, n, x, r, radius, xNorm2, rNorm2, &tlo, &tr);
      if (abspAp <= gam * pNorm * k_nrm2<T>(c, n, q)) {       // pᴴAp ≃ 0
        npcurv = true; stats.indefinite = true; stats.npcCount = 1;
        k_copy<T>(c, n, npc_dir, p);
        if (abspr <= gam * pNorm * rNorm) {                   // pᴴr ≃ 0: p := r
          p = r; q = Ar;
          if (rho > 0) alpha = std::min(tr, rNorm2 / rho);
          else { alpha = tr; if (iter > 0) { stats.npcCount = 2; k_copy<T>(c, n, npc_dir, r); } }
        } else {
          alpha = descent ? t1 : t2;
          if (rho > 0) tr = std::min(tr, rNorm2 / rho);
          const T Delta = -alpha * pr + tr * rNorm2 - tr * tr * rho / 2;
          if (Delta > 0) { p = r; q = Ar; alpha = tr; }
        }
      } else if (pAp > 0 && rho > 0) {
        alpha = rho / k_dot<T>(c, n, q, Mq);
        if (alpha >= t1) { alpha = t1; on_boundary = true; }
      } else if (pAp > 0 && rho < 0) {
        npcurv = true; stats.indefinite = true; stats.npcCount = 1;
        k_copy<T>(c, n, npc_dir, r);
        alpha = descent ? std::min(t1, pr / pAp) : std::max(t2, pr / pAp);
        const T Delta = -alpha * pr + tr * rNorm2 + (alpha * alpha * pAp - tr * tr * rho) / 2;
        if (Delta > 0) { p = r; q = Ar; alpha = tr; }
      } else if (pAp < 0 && rho > 0) {
        npcurv = true; stats.indefinite = true; stats.npcCount = 1;
        k_copy<T>(c, n, npc_dir, p);
        alpha = descent ? t1 : t2;
        tr = std::min(tr, rNorm2 / rho);
        const T Delta = -alpha * pr + tr * rNorm2 + (alpha * alpha * pAp - tr * tr * rho) / 2;
        if (Delta > 0) { p = r; q = Ar; alpha = tr; }
      } else if (pAp < 0 && rho < 0) {
        npcurv = true; stats.indefinite = true; stats.npcCount = 2;
        k_copy<T>(c, n, npc_dir, r);
        alpha = descent ? t1 : t2;
        const T Delta = -alpha * pr + tr * rNorm2 + (alpha * alpha * pAp - tr * tr * rho) / 2;
        if (Delta > 0) { p = r; q = Ar; alpha = tr; }
      }
      // (when the branches above rebind p := r, q := Ar, `Mq` keeps naming the ORIGINAL q array, as in the reference
      //  where Mq was bound once at cr.jl:155; the rebinding always ends the solve in this iteration)
    } else if (radius == 0) {
      alpha = rho / (fusedC && have_qq ? qq : k_dot<T>(c, n, q, Mq));
    }
    T rAr_fused = 0;
    if (fusedC) {
      // grouped passes (fused_phases.cu): x, r updates with ||x||, ||r||; then Ar = A r with ||Ar||, <r, Ar>
      T xx, ArAr;
      cr_fused_step<T>(ws, *A.csr, alpha, &xx, &rNorm2, &ArAr, &rAr_fused);
      xNorm = std::sqrt(xx); rNorm = std::sqrt(rNorm2); ArNorm = std::sqrt(ArAr);
      if (history) { stats.residuals.push_back(rNorm); stats.Aresiduals.push_back(ArNorm); }
    } else {
      k_axpy<T>(c, n, alpha, p, x);
      xNorm = k_nrm2<T>(c, n, x);
      if (radius > 0 && std::fabs(xNorm - radius) <= sqeps * std::max(std::fabs(xNorm), std::fabs(radius))) on_boundary = true;   // xNorm ≈ radius
      k_axpy<T>(c, n, -alpha, Mq, r);
      if (MisI) { rNorm2 = k_dot<T>(c, n, r, r); rNorm = std::sqrt(rNorm2); }
      else {
        const T omega = std::sqrt(alpha) * std::sqrt(rho);
        rNorm = std::sqrt(std::fabs(rNorm + omega)) * std::sqrt(std::fabs(rNorm - omega));
        rNorm2 = rNorm * rNorm;
    }
    if (history) stats.residuals.push_back(rNorm);
    op_apply(c, A, r, Ar);
    ArNorm = k_nrm2<T>(c, n, Ar);
    if (history) stats.Aresiduals.push_back(ArNorm);
    }
    iter = iter + 1;
    if (kdisplay(iter, o.verbose)) {
      mquad = mquad - alpha * pr + alpha * alpha * pAp / 2;
      printf("%5d  %8.1e  %8.1e  %8.1e  %.2fs\n", iter, (double)xNorm, (double)rNorm, (double)mquad, now_seconds() - start_time);
    }
    const bool resid_decrease_mach = (rNorm + T(1) <= T(1));
    if (o.callback) { c.sync(); stats.niter = iter; user_exit = o.callback(&ws, o.callback_user) != 0; }
    const bool resid_decrease = (rNorm <= eps_tol) || resid_decrease_mach;
    solved = resid_decrease || npcurv || on_boundary;
    tired = iter >= itmax;
    overtimed = (now_seconds() - start_time) > o.timemax;
    agree_exit(ws, o, user_exit, overtimed);      // row-partitioned: same decision on every rank
    if (solved || tired || user_exit || overtimed) continue;
    const T rhobar = rho;
    rho = fusedC ? rAr_fused : k_dot<T>(c, n, r, Ar);
    const T beta = rho / rhobar;
    if (fusedC) { qq = cr_fused_directions<T>(ws, beta); have_qq = true; }   // p, q updates + ||q||^2 for the next alpha
    else {
      k_axpby<T>(c, n, T(1), r, beta, p);
      k_axpby<T>(c, n, T(1), Ar, beta, q);
    }
    pNorm2 = rNorm2 + 2 * beta * pr - 2 * beta * alpha * pAp + beta * beta * pNorm2;
    if (pNorm2 > sqeps) pNorm = std::sqrt(pNorm2);
    else if (std::fabs(pNorm2) <= sqeps) pNorm = T(0);
    else {
      stats.niter = iter; stats.solved = solved; stats.inconsistent = false;
      stats.timer = now_seconds() - start_time;
      stats.status = "solver encountered numerical issues";
      if (warm_start) k_axpy<T>(c, n, T(1), dx, x);
      ws.warm_start = false;
      c.sync();
      return;
    }
    pr = rNorm2 + beta * pr - beta * alpha * pAp;
    abspr = std::fabs(pr);
    pAp = rho + beta * beta * pAp;
    abspAp = std::fabs(pAp);
    descent = pr > 0;
  }
  if (o.verbose > 0) printf("\n");
  if (tired) status = "maximum number of iterations exceeded";
  if (solved) status = "solution good enough given atol and rtol";
  if (user_exit) status = "user-requested exit";
  if (overtimed) status = "time limit exceeded";
  if (npcurv) status = "nonpositive curvature";
  if (on_boundary) status = "on trust-region boundary";
  if (warm_start) k_axpy<T>(c, n, T(1), dx, x);
  ws.warm_start = false;
  c.sync();
  stats.niter = iter; stats.solved = solved; stats.inconsistent = false;
  stats.timer = now_seconds() - start_time;
  stats.status = status;
}

#define INST(T)                                                                                                          \
  template void cgs_solve<T>(Workspace<T>&, const LinOp<T>&, const T*, const T*, const LinOp<T>&, const LinOp<T>&, const SolveOpts&); \
  template void cg_lanczos_solve<T>(Workspace<T>&, const LinOp<T>&, const T*, const LinOp<T>&, const SolveOpts&);        \
  template void fom_solve<T>(Workspace<T>&, const LinOp<T>&, const T*, const LinOp<T>&, const LinOp<T>&, const SolveOpts&); \
  template void fgmres_solve<T>(Workspace<T>&, const LinOp<T>&, const T*, const LinOp<T>&, const LinOp<T>&, const SolveOpts&); \
  template void dqgmres_solve<T>(Workspace<T>&, const LinOp<T>&, const T*, const LinOp<T>&, const LinOp<T>&, const SolveOpts&); \
  template void diom_solve<T>(Workspace<T>&, const LinOp<T>&, const T*, const LinOp<T>&, const LinOp<T>&, const SolveOpts&); \
  template void cr_solve<T>(Workspace<T>&, const LinOp<T>&, const T*, const LinOp<T>&, const SolveOpts&);
INST(double)
INST(float)
#undef INST

}  // namespace kb
