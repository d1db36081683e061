// dist.cuh -- cross-GPU pieces of the row-partitioned solve (one process per
// GPU; peers' buffers are mapped with CUDA IPC over NVLink 5 / NVSwitch).
//
// Two mechanisms, both executed INSIDE the compute kernels (no separate
// communication launch, no host involvement):
//
//  * halo gather: the SpMV reads the few x entries owned by other ranks with
//    plain loads from the peers' vectors (P2P over NVLink), row by row, while
//    the interior rows stream from local HBM -- the exchange overlaps the math.
//  * scalar all-reduce: the CTA that finalises a grid-wide dot product writes
//    its rank's partial into every peer's mailbox (P2P stores + system fence +
//    sequence flag) and spins on its own mailbox until all ranks have arrived;
//    every rank sums the same values in rank order, so alpha/beta/stop flags
//    are bit-identical everywhere.  The all-reduce doubles as the inter-GPU
//    barrier that orders halo reads against the peers' vector updates.
//
// The reference has no distributed code; its documentation recipe
// (docs/src/custom_workspaces.md:464-637) does local dot + MPI.Allreduce and a
// user-written distributed mul! -- this is the same decomposition.
#pragma once
#include <cuda_runtime.h>

namespace kb {

constexpr int kMaxRanks = 8;

struct DistComm {
  int rank, world;
  int error;                                   // set on spin timeout
  int pad;
  unsigned long long seq;                      // reductions completed so far (device-resident, stream-ordered)
  double* mail_val[kMaxRanks];                 // [dst rank] -> that rank's value mailbox  [2][kMaxRanks]
  unsigned long long* mail_seq[kMaxRanks];     // [dst rank] -> that rank's flag mailbox   [2][kMaxRanks]
};

// Halo description of a row block: column indices >= nloc refer to entries
// owned by other ranks; entry h lives at offset src_off[h] of rank src_rank[h].
struct HaloMap {
  int nloc;
  int nhalo;
  const int* src_rank;
  const int* src_off;
};

// Push plan: when the entries a peer needs from this rank form contiguous row ranges (slab partitions of
// banded matrices), the kernels that PRODUCE those entries store them straight into the peer's halo buffer
// (posted P2P writes, no round trip) and the consumers gather halo columns from their LOCAL halo buffer.
// Without a push plan the consumers pull halo entries from the owners' vectors (fine-grained P2P loads:
// measured ~60 GB/s effective, 25 us per 46k-entry plane pair).
constexpr int kMaxPushRanges = 4;
struct PushRange { int start, count, peer, slot; };   // local rows [start, start+count) -> peer's halo slots [slot, ...)
template <class T>
struct PushPlan {
  int nranges;                       // 0: pull mode
  PushRange rg[kMaxPushRanges];
  T* dst[kMaxPushRanges];            // peer halo buffer (for the vector being produced) per range
  // Returns true if `row` was sent somewhere: only those threads need the system-scope fence afterwards.
  __device__ __forceinline__ bool operator()(int row, T v) const {
    bool sent = false;
    for (int q = 0; q < nranges; q++) {
      const unsigned d = (unsigned)(row - rg[q].start);
      if (d < (unsigned)rg[q].count) { dst[q][rg[q].slot + d] = v; sent = true; }
    }
    return sent;
  }
};

// General halo exchange for y = A x on ANY workspace vector (all four solvers, primitive and fused-phase
// paths): every rank stores the entries of x its peers need into the peers' x-halo buffer (send list built from
// the halo maps at setup), then the ranks meet in the in-kernel barrier; the SpMV that follows gathers column
// j >= nloc from the local halo buffer.  Two halo sections alternate by exchange parity, and every solver has at
// least one global reduction between two products, so a fast rank can never overwrite a section a slow rank is
// still reading.  (The fused CG keeps its own cheaper scheme: it needs no extra launch.)
struct DistExchange {
  int nsend;                   // entries this rank sends per exchange
  const int* send_row;         // [nsend] local row
  const int* send_peer;        // [nsend] destination rank
  const int* send_slot;        // [nsend] slot in the destination's halo
  void* xhalo_peer[kMaxRanks]; // every rank's x-halo buffer: 2 sections of nhalo_peer[k] elements
  int nhalo_peer[kMaxRanks];
  void* xhalo;                 // this rank's x-halo buffer
  int nhalo, nloc;
  unsigned long long count;    // exchanges done so far (host-side mirror decides the parity)
};

__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Sum `local` over all ranks.  Called by ONE thread per rank (the thread that
// finalises the local grid reduction), the same number of times on every rank.
// Slot parity alternates per reduction: a rank can only be one reduction ahead
// of any peer (it needs the peer's value to finish), so two slots suffice.
__device__ __forceinline__ double dist_allreduce_sum(DistComm* c, double local) {
  const unsigned long long q = c->seq + 1;
  const int base = (int)(q & 1) * kMaxRanks;
  for (int d = 0; d < c->world; d++) st_relaxed_sys(c->mail_val[d] + base + c->rank, local);
  __threadfence_system();
  for (int d = 0; d < c->world; d++) st_relaxed_sys(c->mail_seq[d] + base + c->rank, q);
  double sum = 0.0;
  const unsigned long long* myseq = c->mail_seq[c->rank] + base;
  const double* myval = c->mail_val[c->rank] + base;
  const long long t0 = clock64();
  for (int s = 0; s < c->world; s++) {
    while (ld_acquire_sys(myseq + s) != q) {
      if (clock64() - t0 > 20000000000LL) { c->error = 1; return nan(""); }   // ~10 s: a peer died
    }
    sum += ld_relaxed_sys(myval + s);
  }
  c->seq = q;
  return sum;
}

template <class T>
__device__ __forceinline__ T dist_reduce(DistComm* c, T local) {
  return c ? (T)dist_allreduce_sum(c, (double)local) : local;
}

}  // namespace kb
