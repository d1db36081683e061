// dist.cuh -- cross-GPU pieces of the row-partitioned solve (one process per
// GPU; peers' buffers are mapped with CUDA IPC over NVLink 5 / NVSwitch).
//
// Two mechanisms, both executed INSIDE the compute kernels (no separate
// communication launch, no host involvement):
//
//  * halo: the persistent CG kernel STAGES the x entries owned by other ranks
//    into the tails of its local vectors (coalesced system-scope loads from
//    the peers' vectors, one NVLink round trip per iteration, hidden behind
//    the interior tiles); the two-launch CG kernels pull them nonzero by
//    nonzero; every other product is preceded by halo_exchange_kernel, which
//    pushes the entries into the peers' x-halo buffers (spmv.cu).
//  * scalar all-reduce: the CTA that finalises a grid-wide dot product sends
//    its rank's partial to every peer's mailbox -- value and sequence number
//    in ONE 8-byte store per half (no fence between data and flag) -- and
//    polls its own mailbox until all ranks have arrived; every rank sums the
//    same values in rank order, so alpha/beta/stop flags are bit-identical
//    everywhere.  A full warp does it in parallel in the persistent kernel.
//    The all-reduce doubles as the inter-GPU barrier that orders halo reads
//    against the peers' vector updates.
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
  int error;                                   // sticky: set on spin timeout -- the communicator is DEAD from then on
  int pad;
  unsigned long long seq;                      // reductions completed so far (device-resident, stream-ordered)
  long long timeout_cycles;                    // spin budget of one reduction (KB200_DIST_TIMEOUT_S, default 30 s)
  // [dst rank] -> that rank's mailbox: u64 words [2 parities][kMaxRanks sources][2 halves].  A word carries
  // 32 bits of the value in its low half and the low 32 bits of the reduction's sequence number in its high
  // half (the LL idea of NCCL): value and flag travel in ONE 8-byte store, so no fence separates them.
  unsigned long long* mail[kMaxRanks];
};
constexpr int kMailWords = 2 * kMaxRanks * 2;

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

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Sum `local` over all ranks.  Called by ONE thread per rank (the thread that
// finalises the local grid reduction), the same number of times on every rank.
// Slot parity alternates per reduction: a rank can only be one reduction ahead
// of any peer (it needs the peer's value to finish), so two slots suffice.
// The leading system fence publishes everything this rank wrote before the
// reduction (the peers read halo entries of those vectors afterwards); the
// trailing one orders the peers' data behind the flags just observed.
// After a timeout the communicator is dead: every later reduction returns NaN
// at once (no stale mailbox entry is ever produced) and the host raises.
template <class ACC = double>
__device__ __forceinline__ double dist_allreduce_sum(DistComm* c, double local) {
  if (c->error) return nan("");
  const unsigned long long q = c->seq + 1;
  const unsigned long long tag = (q & 0xffffffffull) << 32;
  const int base = ((int)(q & 1) * kMaxRanks + c->rank) * 2;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(local);
  const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
  __threadfence_system();
  for (int d = 0; d < c->world; d++) {
    st_relaxed_sys(c->mail[d] + base, w0);
    st_relaxed_sys(c->mail[d] + base + 1, w1);
  }
  ACC sum = ACC(0);          // "reduce in T" (SURVEY.md 8e): Float32 solves add the ranks' partials in Float32
  const unsigned long long* mine = c->mail[c->rank] + (int)(q & 1) * kMaxRanks * 2;
  const long long t0 = clock64();
  for (int s = 0; s < c->world; s++) {
    unsigned long long a, b;
    for (;;) {
      a = ld_relaxed_sys_u64(mine + 2 * s); b = ld_relaxed_sys_u64(mine + 2 * s + 1);
      if ((a >> 32) == (tag >> 32) && (b >> 32) == (tag >> 32)) break;
      if (clock64() - t0 > c->timeout_cycles) { c->error = 1; return nan(""); }
    }
    sum += (ACC)__longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
  }
  __threadfence_system();
  c->seq = q;
  return (double)sum;
}

// The same reduction (same mailbox protocol, interoperable call by call) executed by a FULL WARP: lane d sends to
// peer d and polls source d, so the 8 stores and the 8 polls of an 8-GPU box proceed in parallel instead of one
// after the other.  `local` is taken from lane 0; every lane returns the sum (added in rank order).
template <class ACC = double>
__device__ __forceinline__ double dist_allreduce_sum_warp(DistComm* c, double local) {
  const int lane = threadIdx.x & 31;
  local = __shfl_sync(0xffffffffu, local, 0);
  // All loads of the communicator are issued TOGETHER (this runs on the critical path of every barrier of the
  // persistent CG kernel): the constants through the read-only path, lane d fetching the mailbox pointer of peer d,
  // the two words that change (error, seq) as volatile loads -- one memory round trip instead of five dependent ones.
  const int world = __ldg(&c->world), rank = __ldg(&c->rank);
  const long long budget = __ldg(&c->timeout_cycles);
  const unsigned long long box_lane = __ldg(reinterpret_cast<const unsigned long long*>(&c->mail[lane & (kMaxRanks - 1)]));
  const int err = *(volatile int*)&c->error;
  const unsigned long long q = *(volatile unsigned long long*)&c->seq + 1;
  if (err) return nan("");
  const unsigned long long tag = (q & 0xffffffffull) << 32;
  const int base = ((int)(q & 1) * kMaxRanks + rank) * 2;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(local);
  unsigned long long* const peer_box = reinterpret_cast<unsigned long long*>(box_lane);
  const unsigned long long* const my_box = reinterpret_cast<const unsigned long long*>(__shfl_sync(0xffffffffu, box_lane, rank));
  __threadfence_system();
  if (lane < world) {
    st_relaxed_sys(peer_box + base, tag | (bits & 0xffffffffull));
    st_relaxed_sys(peer_box + base + 1, tag | (bits >> 32));
  }
  double v = 0.0;
  bool dead = false;
  if (lane < world) {
    const unsigned long long* src = my_box + ((int)(q & 1) * kMaxRanks + lane) * 2;
    const long long t0 = clock64();
    unsigned long long a, b;
    for (;;) {
      a = ld_relaxed_sys_u64(src); b = ld_relaxed_sys_u64(src + 1);
      if ((a >> 32) == (tag >> 32) && (b >> 32) == (tag >> 32)) break;
      if (clock64() - t0 > budget) { dead = true; break; }
    }
    v = __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
  }
  __threadfence_system();
  if (__any_sync(0xffffffffu, dead)) {
    if (lane == 0) c->error = 1;
    return nan("");
  }
  ACC sum = ACC(0);
  for (int s = 0; s < world; s++) sum += (ACC)__shfl_sync(0xffffffffu, v, s);
  __syncwarp();
  if (lane == 0) c->seq = q;
  return (double)sum;
}

// Partials travel as doubles (exact for Float32) and are added in T, in rank order, identically on every rank.
template <class T>
__device__ __forceinline__ T dist_reduce(DistComm* c, T local) {
  return c ? (T)dist_allreduce_sum<T>(c, (double)local) : local;
}

}  // namespace kb
