// solver_common.h -- small host helpers shared by the solver drivers (solvers.cu, siblings.cu).
#pragma once
#include <cmath>
#include <limits>

#include "kb_internal.h"

namespace kb {

template <class T> static inline T eps_of() { return std::numeric_limits<T>::epsilon(); }
template <class T> static inline T tol_of(double t) { return t < 0 ? std::sqrt(eps_of<T>()) : (T)t; }

// allocate_if (src/krylov_utils.jl:281-288)
template <class T> static inline void allocate_if(bool cond, Workspace<T>& ws, T*& v) {
  const double t0 = now_seconds();
  if (cond && !v) v = dev_alloc<T>((size_t)ws.n);
  ws.stats.allocation_timer += now_seconds() - t0;
}

// Row-partitioned solves: exit decisions taken from a per-rank clock or callback are OR-ed over the ranks (one extra
// tiny launch per iteration, only when a callback or a finite timemax is in play).
template <class T> static inline void agree_exit(Workspace<T>& ws, const SolveOpts& o, bool& user_exit, bool& overtimed) {
  if (ws.dist.world > 1 && (o.callback != nullptr || o.timemax < 1e300)) dist_agree_on_exit(ws.ctx, user_exit, overtimed);
}

static inline bool kdisplay(int iter, int verbose) { return verbose > 0 && iter % verbose == 0; }

// default itmax = 2n of the GLOBAL system (row-partitioned workspaces hold a slice)
template <class T> static inline int default_itmax(const Workspace<T>& ws, int itmax) {
  if (itmax != 0) return itmax;
  const long long two_n = 2LL * (ws.dist.world > 1 ? ws.dist.nglobal : (long long)ws.n);
  return two_n > 2147483647LL ? 2147483647 : (int)two_n;
}

// sym_givens, real case (src/krylov_utils.jl:21-51)
template <class T> static inline void sym_givens(T a, T b, T* c, T* s, T* rho) {
  const T sa = (T)((a > 0) - (a < 0)), sb = (T)((b > 0) - (b < 0));
  if (b == T(0)) { *c = sa + (T)(a == T(0)); *s = T(0); *rho = std::fabs(a); }
  else if (a == T(0)) { *c = T(0); *s = sb; *rho = std::fabs(b); }
  else if (std::fabs(b) > std::fabs(a)) {
    const T t = a / b;
    *s = sb / std::sqrt(T(1) + t * t); *c = *s * t; *rho = b / *s;
  } else {
    const T t = b / a;
    *c = sa / std::sqrt(T(1) + t * t); *s = *c * t; *rho = a / *c;
  }
}

// roots_quadratic (src/krylov_utils.jl:110-152); returns nonzero where the reference raises.
template <class T> static inline int roots_quadratic(T q2, T q1, T q0, int nitref, T* r1, T* r2) {
  T root1, root2;
  if (q2 == T(0)) {
    T root;
    if (q1 == T(0)) { if (q0 != T(0)) return 1; root = T(0); }
    else root = -q0 / q1;
    *r1 = root; *r2 = root;
    return 0;
  }
  const T rhs = std::sqrt(eps_of<T>()) * q1 * q1;
  if (std::fabs(q0 * q2) > rhs) {
    const T rho = q1 * q1 - 4 * q2 * q0;
    if (rho < 0) return 1;
    const T d = -(q1 + std::copysign(std::sqrt(rho), q1)) / 2;
    root1 = d / q2; root2 = q0 / d;
  } else {
    root1 = -q1 / q2; root2 = T(0);
  }
  for (int it = 0; it < nitref; it++) {
    const T q = (q2 * root1 + q1) * root1 + q0, dq = 2 * q2 * root1 + q1;
    if (dq == T(0)) continue;
    root1 = root1 - q / dq;
  }
  for (int it = 0; it < nitref; it++) {
    const T q = (q2 * root2 + q1) * root2 + q0, dq = 2 * q2 * root2 + q1;
    if (dq == T(0)) continue;
    root2 = root2 - q / dq;
  }
  *r1 = root1; *r2 = root2;
  return 0;
}

}  // namespace kb
