/*
 * krylov_oracle_siblings.h -- TEST INFRASTRUCTURE ONLY (same status as
 * krylov_oracle_impl.h, which must be included first).
 *
 * Sibling solvers of SURVEY.md section 8(f)-3: they reuse the hot-path kernels
 * unchanged, so the product runs them through the same primitives.  Literal
 * restatements, one statement per reference line:
 *   cgs!         src/cgs.jl:125-282
 *   cg_lanczos!  src/cg_lanczos.jl:110-264
 *   fom!         src/fom.jl:121-368
 *   fgmres!      src/fgmres.jl:128-388
 * Parity pinning: checked against the properties the reference's own tests
 * assert (test/test_cgs.jl, test_cg_lanczos.jl, test_fom.jl, test_fgmres.jl:
 * residual <= atol + rtol*||b||, stats.solved, status strings) in
 * tests/test_oracle_kat.py; no Julia runtime exists here to compare iterates,
 * so iterate-level parity is "unpinned" exactly as for the four main solvers.
 */
#define PUSH(arr, cnt, v) do { if ((arr) && (cnt) < o->hist_cap) (arr)[(cnt)] = (v); (cnt)++; } while (0)

/* ============================ cgs!  (src/cgs.jl:125-282) ============================ */
int SUF(oracle_cgs)(int n, const int *rowptr, const int *colind, const REAL *val,
                    const REAL *b, const REAL *c_in, const REAL *x0,
                    const REAL *Mdiag, const REAL *Ndiag, const oracle_opts *o,
                    REAL *x, REAL *residuals, oracle_stats *st) {
  SUF(csr) A = {n, rowptr, colind, val};
  memset(st, 0, sizeof(*st));
  set_status(st, "unknown");
  int history = o->history, ldiv = o->ldiv, warm_start = (x0 != NULL);
  int MisI = (Mdiag == NULL), NisI = (Ndiag == NULL);
  const REAL *c = c_in ? c_in : b;                                /* cgs.jl:105 */
  REAL atol = SUF(tol)(o->atol), rtol = SUF(tol)(o->rtol);
  int itmax = o->itmax;
  size_t nb = sizeof(REAL) * (size_t)n;
  REAL *r = malloc(nb), *u = malloc(nb), *p = malloc(nb), *q = malloc(nb), *ts = malloc(nb);
  REAL *vw = MisI ? NULL : malloc(nb), *yz = NisI ? NULL : malloc(nb);
  REAL *t = ts, *s = ts;                                          /* cgs.jl:150-155 */
  REAL *v = MisI ? t : vw, *w = MisI ? s : vw;
  REAL *y = NisI ? p : yz, *z = NisI ? u : yz;
  REAL *r0 = MisI ? r : ts;

  if (warm_start) { SUF(spmv)(&A, x0, r0); SUF(kaxpby)(n, 1, b, -1, r0); }
  else SUF(kcopy)(n, r0, b);
  SUF(kfill)(n, x, 0);
  if (!MisI) SUF(diagmul)(n, r, Mdiag, r0, ldiv);
  REAL rNorm = SUF(knorm)(n, r);
  if (history) PUSH(residuals, st->nres, rNorm);
  if (rNorm == 0) {
    st->niter = 0; st->solved = 1; st->inconsistent = 0;
    set_status(st, "x is a zero-residual solution");
    if (warm_start) SUF(kaxpy)(n, 1, x0, x);
    goto done;
  }
  REAL rho = SUF(kdot)(n, c, r);                                  /* cgs.jl:181 */
  if (rho == 0) {
    st->niter = 0; st->solved = 0; st->inconsistent = 0;
    set_status(st, "Breakdown b\xe1\xb4\xb4" "c = 0");
    if (warm_start) SUF(kaxpy)(n, 1, x0, x);
    goto done;
  }
  int iter = 0;
  if (itmax == 0) itmax = 2 * n;
  REAL eps_ = atol + rtol * rNorm;
  SUF(kcopy)(n, u, r);
  SUF(kcopy)(n, p, r);
  SUF(kfill)(n, q, 0);
  int solved = rNorm <= eps_, tired = iter >= itmax, breakdown = 0;
  while (!(solved || tired || breakdown)) {                       /* cgs.jl:210-252 */
    if (!NisI) SUF(diagmul)(n, y, Ndiag, p, ldiv);
    SUF(spmv)(&A, y, t);
    if (!MisI) SUF(diagmul)(n, v, Mdiag, t, ldiv);
    REAL sigma = SUF(kdot)(n, c, v);
    REAL alpha = rho / sigma;
    SUF(kcopy)(n, q, u);
    SUF(kaxpy)(n, -alpha, v, q);
    SUF(kaxpy)(n, 1, q, u);
    if (!NisI) SUF(diagmul)(n, z, Ndiag, u, ldiv);
    SUF(kaxpy)(n, alpha, z, x);
    SUF(spmv)(&A, z, s);
    if (!MisI) SUF(diagmul)(n, w, Mdiag, s, ldiv);
    SUF(kaxpy)(n, -alpha, w, r);
    REAL rho_next = SUF(kdot)(n, c, r);
    REAL beta = rho_next / rho;
    SUF(kcopy)(n, u, r);
    SUF(kaxpy)(n, beta, q, u);
    SUF(kaxpby)(n, 1, q, beta, p);
    SUF(kaxpby)(n, 1, u, beta, p);
    rho = rho_next;
    iter = iter + 1;
    rNorm = SUF(knorm)(n, r);
    if (history) PUSH(residuals, st->nres, rNorm);
    int resid_decrease_mach = (rNorm + (REAL)1 <= (REAL)1);
    solved = (rNorm <= eps_) || resid_decrease_mach;
    tired = iter >= itmax;
    breakdown = (alpha == 0 || isnan(alpha));
  }
  if (tired) set_status(st, "maximum number of iterations exceeded");
  if (breakdown) set_status(st, "breakdown \xce\xb1\xe2\x82\x96 == 0");
  if (solved) set_status(st, "solution good enough given atol and rtol");
  if (warm_start) SUF(kaxpy)(n, 1, x0, x);
  st->niter = iter; st->solved = solved; st->inconsistent = 0;
done:
  free(r); free(u); free(p); free(q); free(ts); free(vw); free(yz);
  return 0;
}

/* ======================= cg_lanczos!  (src/cg_lanczos.jl:110-264) ======================= */
/* LanczosStats: residuals, indefinite, Anorm (returned through *Anorm), Acond stays NaN. */
int SUF(oracle_cg_lanczos)(int n, const int *rowptr, const int *colind, const REAL *val,
                           const REAL *b, const REAL *x0, const REAL *Mdiag, int check_curvature,
                           const oracle_opts *o, REAL *x, REAL *residuals, REAL *Anorm, oracle_stats *st) {
  SUF(csr) A = {n, rowptr, colind, val};
  memset(st, 0, sizeof(*st));
  set_status(st, "unknown");
  int history = o->history, ldiv = o->ldiv, warm_start = (x0 != NULL);
  int MisI = (Mdiag == NULL);
  REAL atol = SUF(tol)(o->atol), rtol = SUF(tol)(o->rtol);
  int itmax = o->itmax;
  size_t nb = sizeof(REAL) * (size_t)n;
  REAL *Mv = malloc(nb), *Mv_prev = malloc(nb), *p = malloc(nb), *Mv_next = malloc(nb);
  REAL *vbuf = MisI ? NULL : malloc(nb);
  REAL *v = MisI ? Mv : vbuf;                                     /* cg_lanczos.jl:138 */
  *Anorm = (REAL)NAN;

  SUF(kfill)(n, x, 0);
  if (warm_start) { SUF(spmv)(&A, x0, Mv); SUF(kaxpby)(n, 1, b, -1, Mv); }
  else SUF(kcopy)(n, Mv, b);
  if (!MisI) SUF(diagmul)(n, v, Mdiag, Mv, ldiv);
  REAL beta = (v == Mv) ? SUF(knorm)(n, v) : SQRT(SUF(kdot)(n, v, Mv));   /* knorm_elliptic, krylov_utils.jl:319 */
  REAL sigma = beta;
  REAL rNorm = sigma;
  if (history) PUSH(residuals, st->nres, rNorm);
  if (beta == 0) {
    st->niter = 0; st->solved = 1; *Anorm = 0; st->indefinite = 0;
    set_status(st, "x is a zero-residual solution");
    if (warm_start) SUF(kaxpy)(n, 1, x0, x);
    goto done;
  }
  SUF(kcopy)(n, p, v);
  SUF(kdiv)(n, v, beta);
  if (!MisI) SUF(kdiv)(n, Mv, beta);
  SUF(kcopy)(n, Mv_prev, Mv);
  int iter = 0;
  if (itmax == 0) itmax = 2 * n;
  REAL omega = 0, gamma = 1, Anorm2 = 0, beta_prev = 0;
  REAL eps_ = atol + rtol * rNorm;
  int indefinite = 0, solved = rNorm <= eps_, tired = iter >= itmax;
  while (!(solved || tired || (check_curvature && indefinite))) { /* cg_lanczos.jl:190-236 */
    SUF(spmv)(&A, v, Mv_next);
    REAL delta = SUF(kdot)(n, v, Mv_next);
    gamma = (REAL)1 / (delta - omega / gamma);
    indefinite |= (gamma <= 0);
    if (check_curvature && indefinite) continue;
    SUF(kaxpy)(n, -delta, Mv, Mv_next);
    if (iter > 0) {
      SUF(kaxpy)(n, -beta, Mv_prev, Mv_next);
      SUF(kcopy)(n, Mv_prev, Mv);
    }
    SUF(kcopy)(n, Mv, Mv_next);
    if (!MisI) SUF(diagmul)(n, v, Mdiag, Mv, ldiv);
    beta = (v == Mv) ? SUF(knorm)(n, v) : SQRT(SUF(kdot)(n, v, Mv));
    SUF(kdiv)(n, v, beta);
    if (!MisI) SUF(kdiv)(n, Mv, beta);
    Anorm2 += beta_prev * beta_prev + beta * beta + delta * delta;
    beta_prev = beta;
    SUF(kaxpy)(n, gamma, p, x);
    omega = beta * gamma;
    sigma = -omega * sigma;
    omega = omega * omega;
    SUF(kaxpby)(n, sigma, v, omega, p);
    rNorm = FABS(sigma);
    if (history) PUSH(residuals, st->nres, rNorm);
    iter = iter + 1;
    int resid_decrease_mach = (rNorm + (REAL)1 <= (REAL)1);
    solved = (rNorm <= eps_) || resid_decrease_mach;
    tired = iter >= itmax;
  }
  if (tired) set_status(st, "maximum number of iterations exceeded");
  if (check_curvature && indefinite) set_status(st, "negative curvature");
  if (solved) set_status(st, "solution good enough given atol and rtol");
  if (warm_start) SUF(kaxpy)(n, 1, x0, x);
  st->niter = iter; st->solved = solved; *Anorm = SQRT(Anorm2); st->indefinite = indefinite;
done:
  free(Mv); free(Mv_prev); free(p); free(Mv_next); free(vbuf);
  return 0;
}

/* ============================ fom!  (src/fom.jl:121-368) ============================ */
int SUF(oracle_fom)(int n, const int *rowptr, const int *colind, const REAL *val,
                    const REAL *b, const REAL *x0, const REAL *Mdiag, const REAL *Ndiag,
                    const oracle_opts *o, REAL *x, REAL *residuals, oracle_stats *st) {
  SUF(csr) A = {n, rowptr, colind, val};
  memset(st, 0, sizeof(*st));
  set_status(st, "unknown");
  int history = o->history, ldiv = o->ldiv, warm_start = (x0 != NULL);
  int restart = o->restart, reorth = o->reorthogonalization;
  int MisI = (Mdiag == NULL), NisI = (Ndiag == NULL);
  REAL atol = SUF(tol)(o->atol), rtol = SUF(tol)(o->rtol);
  int itmax = o->itmax;
  int mem = o->memory == 0 ? 20 : o->memory;
  if (mem > n) mem = n;                                           /* krylov_workspaces.jl:3067 */
  size_t nb = sizeof(REAL) * (size_t)n;
  REAL *w = malloc(nb);
  REAL *qbuf = MisI ? NULL : malloc(nb), *pbuf = NisI ? NULL : malloc(nb);
  REAL *dx = (restart || warm_start) ? calloc(n, sizeof(REAL)) : NULL;
  if (warm_start) SUF(kcopy)(n, dx, x0);
  int vcap = mem, lcap = mem, ucap = mem * (mem + 1) / 2;
  REAL **V = malloc(sizeof(REAL *) * vcap);
  for (int i = 0; i < vcap; i++) V[i] = malloc(nb);
  REAL *l = malloc(sizeof(REAL) * lcap), *zz = malloc(sizeof(REAL) * lcap), *U = malloc(sizeof(REAL) * ucap);
  int llen = mem, ulen = ucap, vlen = mem;                        /* Julia vector lengths (l and z grow together) */
  REAL *q = MisI ? w : qbuf, *r0 = MisI ? w : qbuf;               /* fom.jl:150-152 */
  REAL *xr = restart ? dx : x;

  SUF(kfill)(n, x, 0);
  if (warm_start) {
    SUF(spmv)(&A, dx, w);
    SUF(kaxpby)(n, 1, b, -1, w);
    if (restart) SUF(kaxpy)(n, 1, dx, x);
  } else {
    SUF(kcopy)(n, w, b);
  }
  if (!MisI) SUF(diagmul)(n, r0, Mdiag, w, ldiv);
  REAL beta = SUF(knorm)(n, r0);
  REAL rNorm = beta;
  if (history) PUSH(residuals, st->nres, beta);
  REAL eps_ = atol + rtol * rNorm;
  if (beta == 0) {
    st->niter = 0; st->solved = 1; st->inconsistent = 0;
    set_status(st, "x is a zero-residual solution");
    if (warm_start) SUF(kaxpy)(n, 1, x0, x);
    goto done;
  }
  mem = llen;                                                     /* fom.jl:183 */
  int npass = 0, iter = 0, inner_iter = 0;
  if (itmax == 0) itmax = 2 * n;
  int inner_itmax = itmax;
  REAL btol = POW(EPS, (REAL)0.75);
  int breakdown = 0, solved = rNorm <= eps_, tired = iter >= itmax, inner_tired;

  while (!(solved || tired || breakdown)) {                       /* fom.jl:207 */
    int nr = 0;
    for (int i = 0; i < mem; i++) SUF(kfill)(n, V[i], 0);
    for (int i = 0; i < llen; i++) { l[i] = 0; zz[i] = 0; }
    for (int i = 0; i < ulen; i++) U[i] = 0;
    if (restart) {
      SUF(kfill)(n, xr, 0);
      if (npass >= 1) {
        SUF(spmv)(&A, x, w);
        SUF(kaxpby)(n, 1, b, -1, w);
        if (!MisI) SUF(diagmul)(n, r0, Mdiag, w, ldiv);
      }
    }
    beta = SUF(knorm)(n, r0);
    zz[0] = beta;
    SUF(kdivcopy)(n, V[0], r0, rNorm);                            /* fom.jl:231 */
    npass = npass + 1;
    inner_iter = 0;
    inner_tired = 0;
    while (!(solved || inner_tired || breakdown)) {               /* fom.jl:237 */
      inner_iter = inner_iter + 1;
      if (!restart && (inner_iter > mem)) {                       /* fom.jl:243-251 */
        int newu = ulen + inner_iter;
        if (newu > ucap) { ucap = 2 * newu; U = realloc(U, sizeof(REAL) * ucap); }
        for (int i = ulen; i < newu; i++) U[i] = 0;
        ulen = newu;
        if (llen + 1 > lcap) { lcap = 2 * (llen + 1); l = realloc(l, sizeof(REAL) * lcap); zz = realloc(zz, sizeof(REAL) * lcap); }
        l[llen] = 0; zz[llen] = 0; llen++;
      }
      REAL *pv = V[inner_iter - 1];
      REAL *p = NisI ? pv : pbuf;
      if (!NisI) SUF(diagmul)(n, p, Ndiag, pv, ldiv);
      SUF(spmv)(&A, p, w);
      if (!MisI) SUF(diagmul)(n, q, Mdiag, w, ldiv);
      for (int i = 0; i < inner_iter; i++) {
        U[nr + i] = SUF(kdot)(n, V[i], q);
        SUF(kaxpy)(n, -U[nr + i], V[i], q);
      }
      if (reorth) {
        for (int i = 0; i < inner_iter; i++) {
          REAL Htmp = SUF(kdot)(n, V[i], q);
          U[nr + i] += Htmp;
          SUF(kaxpy)(n, -Htmp, V[i], q);
        }
      }
      REAL Hbis = SUF(knorm)(n, q);
      if (inner_iter >= 2) {                                      /* LU update, fom.jl:274-281 */
        for (int i = 2; i <= inner_iter; i++) U[nr + i - 1] = U[nr + i - 1] - l[i - 2] * U[nr + i - 2];
        zz[inner_iter - 1] = -l[inner_iter - 2] * zz[inner_iter - 2];
      }
      l[inner_iter - 1] = Hbis / U[nr + inner_iter - 1];
      rNorm = Hbis * FABS(zz[inner_iter - 1] / U[nr + inner_iter - 1]);
      if (history) PUSH(residuals, st->nres, rNorm);
      nr = nr + inner_iter;
      int resid_decrease_mach = (rNorm + (REAL)1 <= (REAL)1);
      breakdown = Hbis <= btol;
      solved = (rNorm <= eps_) || resid_decrease_mach;
      {
        int lim = restart ? (mem < inner_itmax ? mem : inner_itmax) : inner_itmax;
        inner_tired = inner_iter >= lim;
      }
      if (!(solved || inner_tired || breakdown)) {
        if (!restart && (inner_iter >= mem)) {
          if (vlen + 1 > vcap) { vcap = 2 * (vlen + 1); V = realloc(V, sizeof(REAL *) * vcap); }
          V[vlen++] = malloc(nb);
        }
        SUF(kdivcopy)(n, V[inner_iter], q, Hbis);
      }
    }
    REAL *y = zz;                                                 /* fom.jl:322-331 */
    for (int i = inner_iter; i >= 1; i--) {
      int pos = nr + i - inner_iter;
      for (int j = inner_iter; j >= i + 1; j--) {
        y[i - 1] = y[i - 1] - U[pos - 1] * y[j - 1];
        pos = pos - j + 1;
      }
      y[i - 1] = y[i - 1] / U[pos - 1];
    }
    for (int i = 0; i < inner_iter; i++) SUF(kaxpy)(n, y[i], V[i], xr);
    if (!NisI) { SUF(kcopy)(n, pbuf, xr); SUF(diagmul)(n, xr, Ndiag, pbuf, ldiv); }
    if (restart) SUF(kaxpy)(n, 1, xr, x);
    inner_itmax = inner_itmax - inner_iter;
    iter = iter + inner_iter;
    tired = iter >= itmax;
  }
  if (tired) set_status(st, "maximum number of iterations exceeded");
  if (breakdown) set_status(st, "inconsistent linear system");
  if (solved) set_status(st, "solution good enough given atol and rtol");
  if (warm_start && !restart) SUF(kaxpy)(n, 1, x0, x);
  st->niter = iter; st->solved = solved; st->inconsistent = !solved && breakdown;
done:
  for (int i = 0; i < vlen; i++) free(V[i]);
  free(V); free(l); free(zz); free(U); free(w); free(qbuf); free(pbuf); free(dx);
  return 0;
}

/* ========================== fgmres!  (src/fgmres.jl:128-388) ========================== */
/* Ndiag == NULL: N === I, and Z[k] <- V[k] by the unconditional mulorldiv! (fgmres.jl:262). */
int SUF(oracle_fgmres)(int n, const int *rowptr, const int *colind, const REAL *val,
                       const REAL *b, const REAL *x0, const REAL *Mdiag, const REAL *Ndiag,
                       const oracle_opts *o, REAL *x, REAL *residuals, oracle_stats *st) {
  SUF(csr) A = {n, rowptr, colind, val};
  memset(st, 0, sizeof(*st));
  set_status(st, "unknown");
  int history = o->history, ldiv = o->ldiv, warm_start = (x0 != NULL);
  int restart = o->restart, reorth = o->reorthogonalization;
  int MisI = (Mdiag == NULL), NisI = (Ndiag == NULL);
  REAL atol = SUF(tol)(o->atol), rtol = SUF(tol)(o->rtol);
  int itmax = o->itmax;
  int mem = o->memory == 0 ? 20 : o->memory;
  if (mem > n) mem = n;                                           /* krylov_workspaces.jl:2985 */
  size_t nb = sizeof(REAL) * (size_t)n;
  REAL *w = malloc(nb);
  REAL *qbuf = MisI ? NULL : malloc(nb);
  REAL *dx = (restart || warm_start) ? calloc(n, sizeof(REAL)) : NULL;
  if (warm_start) SUF(kcopy)(n, dx, x0);
  int vcap = mem, zvcap = mem, scap = mem, rcap = mem * (mem + 1) / 2, zcap = mem;
  REAL **V = malloc(sizeof(REAL *) * vcap), **Z = malloc(sizeof(REAL *) * zvcap);
  for (int i = 0; i < vcap; i++) { V[i] = malloc(nb); Z[i] = malloc(nb); }
  REAL *cc = malloc(sizeof(REAL) * scap), *ss = malloc(sizeof(REAL) * scap);
  REAL *zz = malloc(sizeof(REAL) * zcap), *R = malloc(sizeof(REAL) * rcap);
  int clen = mem, rlen = rcap, vlen = mem, zvlen = mem, zlen = mem;
  REAL *q = MisI ? w : qbuf, *r0 = MisI ? w : qbuf;
  REAL *xr = restart ? dx : x;

  SUF(kfill)(n, x, 0);
  if (warm_start) {
    SUF(spmv)(&A, dx, w);
    SUF(kaxpby)(n, 1, b, -1, w);
    if (restart) SUF(kaxpy)(n, 1, dx, x);
  } else {
    SUF(kcopy)(n, w, b);
  }
  if (!MisI) SUF(diagmul)(n, r0, Mdiag, w, ldiv);
  REAL beta = SUF(knorm)(n, r0);
  REAL rNorm = beta;
  if (history) PUSH(residuals, st->nres, beta);
  REAL eps_ = atol + rtol * rNorm;
  if (beta == 0) {
    st->niter = 0; st->solved = 1; st->inconsistent = 0;
    set_status(st, "x is a zero-residual solution");
    if (warm_start) SUF(kaxpy)(n, 1, x0, x);
    goto done;
  }
  mem = clen;
  int npass = 0, iter = 0, inner_iter = 0;
  if (itmax == 0) itmax = 2 * n;
  int inner_itmax = itmax;
  REAL btol = POW(EPS, (REAL)0.75);
  int breakdown = 0, inconsistent = 0, solved = rNorm <= eps_, tired = iter >= itmax, inner_tired;

  while (!(solved || tired || breakdown)) {
    int nr = 0;
    for (int i = 0; i < mem; i++) { SUF(kfill)(n, V[i], 0); SUF(kfill)(n, Z[i], 0); }
    for (int i = 0; i < clen; i++) { ss[i] = 0; cc[i] = 0; }
    for (int i = 0; i < rlen; i++) R[i] = 0;
    for (int i = 0; i < zlen; i++) zz[i] = 0;
    if (restart) {
      SUF(kfill)(n, xr, 0);
      if (npass >= 1) {
        SUF(spmv)(&A, x, w);
        SUF(kaxpby)(n, 1, b, -1, w);
        if (!MisI) SUF(diagmul)(n, r0, Mdiag, w, ldiv);
      }
    }
    beta = SUF(knorm)(n, r0);
    zz[0] = beta;
    SUF(kdivcopy)(n, V[0], r0, rNorm);
    npass = npass + 1;
    inner_iter = 0;
    inner_tired = 0;
    while (!(solved || inner_tired || breakdown)) {
      inner_iter = inner_iter + 1;
      if (!restart && (inner_iter > mem)) {                       /* fgmres.jl:250-259 */
        int newr = rlen + inner_iter;
        if (newr > rcap) { rcap = 2 * newr; R = realloc(R, sizeof(REAL) * rcap); }
        for (int i = rlen; i < newr; i++) R[i] = 0;
        rlen = newr;
        if (clen + 1 > scap) { scap = 2 * (clen + 1); cc = realloc(cc, sizeof(REAL) * scap); ss = realloc(ss, sizeof(REAL) * scap); }
        ss[clen] = 0; cc[clen] = 0; clen++;
        if (zvlen + 1 > zvcap) { zvcap = 2 * (zvlen + 1); Z = realloc(Z, sizeof(REAL *) * zvcap); }
        Z[zvlen++] = malloc(nb);
      }
      REAL *zk = Z[inner_iter - 1];
      if (NisI) SUF(kcopy)(n, zk, V[inner_iter - 1]);             /* mul!(z, I, v) */
      else SUF(diagmul)(n, zk, Ndiag, V[inner_iter - 1], ldiv);
      SUF(spmv)(&A, zk, w);
      if (!MisI) SUF(diagmul)(n, q, Mdiag, w, ldiv);
      for (int i = 0; i < inner_iter; i++) {
        R[nr + i] = SUF(kdot)(n, V[i], q);
        SUF(kaxpy)(n, -R[nr + i], V[i], q);
      }
      if (reorth) {
        for (int i = 0; i < inner_iter; i++) {
          REAL Htmp = SUF(kdot)(n, V[i], q);
          R[nr + i] += Htmp;
          SUF(kaxpy)(n, -Htmp, V[i], q);
        }
      }
      REAL Hbis = SUF(knorm)(n, q);
      for (int i = 0; i < inner_iter - 1; i++) {
        REAL Rtmp = cc[i] * R[nr + i] + ss[i] * R[nr + i + 1];
        R[nr + i + 1] = ss[i] * R[nr + i] - cc[i] * R[nr + i + 1];
        R[nr + i] = Rtmp;
      }
      SUF(oracle_sym_givens)(R[nr + inner_iter - 1], Hbis, &cc[inner_iter - 1], &ss[inner_iter - 1], &R[nr + inner_iter - 1]);
      REAL zeta_next = ss[inner_iter - 1] * zz[inner_iter - 1];
      zz[inner_iter - 1] = cc[inner_iter - 1] * zz[inner_iter - 1];
      rNorm = FABS(zeta_next);
      if (history) PUSH(residuals, st->nres, rNorm);
      nr = nr + inner_iter;
      int resid_decrease_mach = (rNorm + (REAL)1 <= (REAL)1);
      breakdown = Hbis <= btol;
      solved = (rNorm <= eps_) || resid_decrease_mach;
      {
        int lim = restart ? (mem < inner_itmax ? mem : inner_itmax) : inner_itmax;
        inner_tired = inner_iter >= lim;
      }
      if (!(solved || inner_tired || breakdown)) {
        if (!restart && (inner_iter >= mem)) {
          if (vlen + 1 > vcap) { vcap = 2 * (vlen + 1); V = realloc(V, sizeof(REAL *) * vcap); }
          V[vlen++] = malloc(nb);
          if (zlen + 1 > zcap) { zcap = 2 * (zlen + 1); zz = realloc(zz, sizeof(REAL) * zcap); }
          zz[zlen++] = 0;
        }
        SUF(kdivcopy)(n, V[inner_iter], q, Hbis);
        zz[inner_iter] = zeta_next;
      }
    }
    REAL *y = zz;
    for (int i = inner_iter; i >= 1; i--) {
      int pos = nr + i - inner_iter;
      for (int j = inner_iter; j >= i + 1; j--) {
        y[i - 1] = y[i - 1] - R[pos - 1] * y[j - 1];
        pos = pos - j + 1;
      }
      if (FABS(R[pos - 1]) <= btol) { y[i - 1] = 0; inconsistent = 1; }
      else y[i - 1] = y[i - 1] / R[pos - 1];
    }
    for (int i = 0; i < inner_iter; i++) SUF(kaxpy)(n, y[i], Z[i], xr);   /* fgmres.jl:355-357 */
    if (restart) SUF(kaxpy)(n, 1, xr, x);
    inner_itmax = inner_itmax - inner_iter;
    iter = iter + inner_iter;
    tired = iter >= itmax;
  }
  if (tired) set_status(st, "maximum number of iterations exceeded");
  if (solved) set_status(st, "solution good enough given atol and rtol");
  if (inconsistent) set_status(st, "found approximate least-squares solution");
  if (warm_start && !restart) SUF(kaxpy)(n, 1, x0, x);
  st->niter = iter; st->solved = solved; st->inconsistent = inconsistent;
done:
  for (int i = 0; i < vlen; i++) free(V[i]);
  for (int i = 0; i < zvlen; i++) free(Z[i]);
  free(V); free(Z); free(cc); free(ss); free(zz); free(R); free(w); free(qbuf); free(dx);
  return 0;
}

#undef PUSH
