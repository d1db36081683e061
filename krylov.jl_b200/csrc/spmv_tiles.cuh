// spmv_tiles.cuh -- TMA-staged CSR row-tile pipeline (sm_100a).
//
// A persistent CTA walks tiles of kTileRows consecutive rows.  One producer
// warp streams each tile's row-pointer slice, column indices and values from
// HBM into a shared-memory ring with 1-D bulk TMA copies (cp.async.bulk ->
// SASS UBLKCP) that complete on an mbarrier; eight consumer warps wait on the
// barrier and reduce one row per thread from shared memory, gathering x
// through the read-only L1/L2 path.  Consumers hand the stage back through a
// second mbarrier.  The matrix (>= 80 % of the bytes of an SpMV) therefore
// moves as large asynchronous bursts with several KB in flight per SM and no
// register staging, while the irregular x gather stays on LDG.
//
// Row sums accumulate left to right in ascending column order with the product
// rounded before the add -- the order SparseArrays' CSC mul! produces for every
// y[i] -- so y is bit-identical to the sequential CPU oracle.
#pragma once
#include "common.cuh"
#include "kb_internal.h"

namespace kb {

constexpr int kConsumerWarps = kTileRows / 32;            // 8
constexpr int kTileThreads = kTileRows + 32;              // + 1 producer warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D bulk TMA copy global -> shared, completion counted in bytes on `bar`.
// The matrix streams are read once per SpMV: tag them evict-first in L2 so the
// gathered vectors (re-read by neighbouring rows) keep their lines.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// Shared-memory layout of one stage (all offsets multiples of 128 B).
template <class T>
struct TileLayout {
  int cap;  // max nnz per tile
  __host__ __device__ static constexpr size_t align_up(size_t v) { return (v + 127) & ~size_t(127); }
  __host__ __device__ size_t rp_bytes() const { return align_up((kTileRows + 4) * sizeof(int)); }
  __host__ __device__ size_t val_bytes() const { return align_up((size_t)(cap + 16 / sizeof(T)) * sizeof(T)); }
  __host__ __device__ size_t col_bytes() const { return align_up((size_t)(cap + 8) * sizeof(int)); }
  __host__ __device__ size_t stage_bytes() const { return rp_bytes() + val_bytes() + col_bytes(); }
  __host__ __device__ size_t total_bytes(int stages) const { return 128 + (size_t)stages * stage_bytes(); }
};

// Runs the tile pipeline.  Every thread of the CTA must call it (blockDim.x ==
// kTileThreads).  `gather(j)` returns the x value for column j; `row_done(row,
// acc)` receives each finished row sum (consumer threads only, row < n).
template <class T, class Gather, class RowDone>
__device__ __forceinline__ void spmv_tiles_run(const Csr<T>& A, unsigned char* smem, Gather gather, RowDone row_done) {
  const TileLayout<T> L{A.tile_cap};
  const int S = A.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [S]
  uint64_t* empty = full + S;                                // [S]   (S <= 8 -> 128 B header)
  unsigned char* ring = smem + 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int VA = 16 / sizeof(T);                         // values per 16 B

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      const uint64_t pol = l2_evict_first_policy();
      int it = 0;
      for (int t = blockIdx.x; t < A.ntiles; t += gridDim.x, it++) {
        const int s = it % S;
        mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
        unsigned char* st = ring + (size_t)s * L.stage_bytes();
        const int r0 = t * kTileRows;
        const int r1 = min(r0 + kTileRows, A.n);
        const int k0 = __ldg(&A.rowptr[r0]), k1 = __ldg(&A.rowptr[r1]);
        const int k0v = k0 & ~(VA - 1), k1v = (k1 + VA - 1) & ~(VA - 1);
        const int k0c = k0 & ~3, k1c = (k1 + 3) & ~3;
        const unsigned rp_b = (kTileRows + 4) * sizeof(int);
        const unsigned v_b = (unsigned)(k1v - k0v) * sizeof(T);
        const unsigned c_b = (unsigned)(k1c - k0c) * sizeof(int);
        mbar_expect_tx(&full[s], rp_b + v_b + c_b);
        tma_load_1d(st, A.rowptr + r0, rp_b, &full[s], pol);
        if (v_b) tma_load_1d(st + L.rp_bytes(), A.val + k0v, v_b, &full[s], pol);
        if (c_b) tma_load_1d(st + L.rp_bytes() + L.val_bytes(), A.colind + k0c, c_b, &full[s], pol);
      }
    }
  } else {
    // ------------------------------ consumers -----------------------------
    int it = 0;
    for (int t = blockIdx.x; t < A.ntiles; t += gridDim.x, it++) {
      const int s = it % S;
      mbar_wait(&full[s], (it / S) & 1);
      const unsigned char* st = ring + (size_t)s * L.stage_bytes();
      const int* rp = reinterpret_cast<const int*>(st);
      const T* vs = reinterpret_cast<const T*>(st + L.rp_bytes());
      const int* cs = reinterpret_cast<const int*>(st + L.rp_bytes() + L.val_bytes());
      const int row = t * kTileRows + tid;
      if (row < A.n) {
        const int k0 = rp[0];
        const int voff = k0 & ~(VA - 1), coff = k0 & ~3;
        const int kb = rp[tid], ke = rp[tid + 1];
        T acc = T(0);
        int k = kb;
        // two nonzeros per trip: both gathers are issued before either is consumed
        for (; k + 1 < ke; k += 2) {
          const int c0 = cs[k - coff], c1 = cs[k + 1 - coff];
          const T a0 = vs[k - voff], a1 = vs[k + 1 - voff];
          const T x0 = gather(c0), x1 = gather(c1);
          acc = add_rn(acc, mul_rn(a0, x0));
          acc = add_rn(acc, mul_rn(a1, x1));
        }
        if (k < ke) acc = add_rn(acc, mul_rn(vs[k - voff], gather(cs[k - coff])));
        row_done(row, acc);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

}  // namespace kb
