// spmv.cu -- the CSR operator: upload/normalisation, staging plan, and the
// y = A x kernels that stand in for `kmul!(y, A, x)` = mul!(y, A, x)
// (src/krylov_utils.jl:305; call sites cg.jl:196, gmres.jl:257,
// bicgstab.jl:221,228, minres.jl:289).
#include "kb_internal.h"
#include "spmv_tiles.cuh"

#include <vector>

namespace kb {

// ---------------------------------------------------------------------------
// Upload: accept the caller's (rowptr, colind, val) with 0/1-based, 32/64-bit
// indices on host or device; store int32 0-based in padded device arrays.
// The bijection (shift by index_base, narrow to int32) keeps every
// (row, col, val) triplet of the input -- "bit-exact integer indexing".
// ---------------------------------------------------------------------------
template <class I>
__global__ void index_convert_kernel(long long cnt, const I* __restrict__ in, int* __restrict__ out, int base, int* bad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < cnt; i += stride) {
    long long v = (long long)in[i] - base;
    if (v < 0 || v > 2147483647LL) atomicExch(bad, 1);
    out[i] = (int)v;
  }
}

__global__ void fill_int_kernel(int cnt, int* out, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) out[i] = v;
}

template <class T>
void csr_upload(Ctx& c, Csr<T>& A, int n, long long nnz, const void* rowptr, const void* colind, const T* val,
                int index_base, int index_bytes, bool on_device) {
  if (n < 0 || nnz < 0 || nnz > 2147483647LL - 64) throw std::runtime_error("CSR operator: n/nnz out of int32 range");
  if (index_bytes != 4 && index_bytes != 8) throw std::runtime_error("CSR operator: index_bytes must be 4 or 8");
  if (index_base != 0 && index_base != 1) throw std::runtime_error("CSR operator: index_base must be 0 or 1");
  csr_free(A);
  A.n = n; A.nnz = nnz;
  const size_t rp_len = (size_t)n + 1, rp_pad = kTileRows + 16;
  A.rowptr = dev_alloc<int>(rp_len + rp_pad);
  A.colind = dev_alloc<int>((size_t)nnz + 16);
  A.val = dev_alloc<T>((size_t)nnz + 16);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  KB_CUDA(cudaMemcpyAsync(A.val, val, sizeof(T) * (size_t)nnz, kind, c.stream));
  KB_CUDA(cudaMemsetAsync(A.val + nnz, 0, sizeof(T) * 16, c.stream));
  KB_CUDA(cudaMemsetAsync(A.colind + nnz, 0, sizeof(int) * 16, c.stream));
  int* bad = nullptr;
  KB_CUDA(cudaMalloc((void**)&bad, sizeof(int)));
  KB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c.stream));
  void* tmp_rp = nullptr; void* tmp_ci = nullptr;
  if (index_bytes == 4 && index_base == 0) {
    KB_CUDA(cudaMemcpyAsync(A.rowptr, rowptr, sizeof(int) * rp_len, kind, c.stream));
    KB_CUDA(cudaMemcpyAsync(A.colind, colind, sizeof(int) * (size_t)nnz, kind, c.stream));
  } else {
    const void* drp = rowptr; const void* dci = colind;
    if (!on_device) {
      KB_CUDA(cudaMalloc(&tmp_rp, (size_t)index_bytes * rp_len));
      KB_CUDA(cudaMalloc(&tmp_ci, (size_t)index_bytes * (size_t)(nnz ? nnz : 1)));
      KB_CUDA(cudaMemcpyAsync(tmp_rp, rowptr, (size_t)index_bytes * rp_len, cudaMemcpyHostToDevice, c.stream));
      KB_CUDA(cudaMemcpyAsync(tmp_ci, colind, (size_t)index_bytes * (size_t)nnz, cudaMemcpyHostToDevice, c.stream));
      drp = tmp_rp; dci = tmp_ci;
    }
    const int g = sm_count() * 8;
    if (index_bytes == 8) {
      index_convert_kernel<long long><<<g, 256, 0, c.stream>>>((long long)rp_len, (const long long*)drp, A.rowptr, index_base, bad);
      index_convert_kernel<long long><<<g, 256, 0, c.stream>>>(nnz, (const long long*)dci, A.colind, index_base, bad);
    } else {
      index_convert_kernel<int><<<g, 256, 0, c.stream>>>((long long)rp_len, (const int*)drp, A.rowptr, index_base, bad);
      index_convert_kernel<int><<<g, 256, 0, c.stream>>>(nnz, (const int*)dci, A.colind, index_base, bad);
    }
    KB_CUDA(cudaGetLastError());
  }
  // rows past n (read by the last tile's row-pointer slice) are empty
  fill_int_kernel<<<((int)rp_pad + 255) / 256, 256, 0, c.stream>>>((int)rp_pad, A.rowptr + rp_len, (int)nnz);
  KB_CUDA(cudaGetLastError());
  int hbad = 0;
  KB_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  cudaFree(bad);
  if (tmp_rp) cudaFree(tmp_rp);
  if (tmp_ci) cudaFree(tmp_ci);
  if (hbad) { csr_free(A); throw std::runtime_error("CSR operator: index outside int32 range after rebasing"); }
  try { csr_plan(c, A); } catch (...) { csr_free(A); throw; }
}

template <class T> void csr_free(Csr<T>& A) {
  dev_free(A.rowptr); dev_free(A.colind); dev_free(A.val);
  A = Csr<T>();
}

// ---------------------------------------------------------------------------
// Staging plan: largest tile (nnz of kTileRows consecutive rows) and longest
// row decide whether the TMA ring fits, how deep it is, and the grid.
// ---------------------------------------------------------------------------
__global__ void plan_kernel(int n, int ntiles, const int* __restrict__ rowptr,
                            int* out /* [0]=tile_cap [1]=max_row [2]=unsorted [3]=rowptr not monotone [4]=max col [5]=negative col */,
                            const int* __restrict__ colind, long long nnz) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  int cap = 0, mr = 0, uns = 0, bad = 0, mc = -1, neg = 0;
  for (int i = t; i < n; i += stride) {
    const int kb = rowptr[i], ke = rowptr[i + 1];
    // validation (a malformed matrix must be an error, not an out-of-bounds read in the SpMV): row pointers
    // non-decreasing and inside [0, nnz] -- only then are the column indices of the row looked at
    if (kb > ke || kb < 0 || (long long)ke > nnz) { bad = 1; continue; }
    mr = max(mr, ke - kb);
    for (int k = kb; k < ke; k++) { const int cj = colind[k]; mc = max(mc, cj); neg |= cj < 0; }
    // halo columns (index >= n, row-partitioned operators) keep their global position in the row: skip them
    for (int k = kb + 1; k < ke; k++) uns |= (colind[k] <= colind[k - 1]) && colind[k] < n && colind[k - 1] < n;
  }
  if (bad) atomicExch(&out[3], 1);
  __syncthreads();
  for (int i = t; i < ntiles; i += stride) {
    const int r0 = i * kTileRows, r1 = min(r0 + kTileRows, n);
    cap = max(cap, rowptr[r1] - rowptr[r0]);
  }
  atomicMax(&out[0], cap);
  atomicMax(&out[1], mr);
  atomicMax(&out[4], mc);
  if (uns) atomicExch(&out[2], 1);
  if (neg) atomicExch(&out[5], 1);
}

template <class T> void csr_plan(Ctx& c, Csr<T>& A) {
  A.ntiles = (A.n + kTileRows - 1) / kTileRows;
  int* dout = nullptr;
  KB_CUDA(cudaMalloc((void**)&dout, 6 * sizeof(int)));
  KB_CUDA(cudaMemsetAsync(dout, 0, 6 * sizeof(int), c.stream));
  int h[6] = {0, 0, 0, 0, -1, 0};
  int ends[2] = {0, (int)A.nnz};
  if (A.n > 0) {
    KB_CUDA(cudaMemcpyAsync(dout + 4, &h[4], sizeof(int), cudaMemcpyHostToDevice, c.stream));
    plan_kernel<<<sm_count() * 4, 256, 0, c.stream>>>(A.n, A.ntiles, A.rowptr, dout, A.colind, A.nnz);
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaMemcpyAsync(&ends[0], A.rowptr, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    KB_CUDA(cudaMemcpyAsync(&ends[1], A.rowptr + A.n, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  }
  KB_CUDA(cudaMemcpyAsync(h, dout, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  c.sync();
  cudaFree(dout);
  // the reference would throw a BoundsError on such input; here it must not reach the kernels
  if (ends[0] != 0 || (long long)ends[1] != A.nnz || h[3])
    throw std::runtime_error("CSR operator: row pointers must start at 0 (after rebasing), be non-decreasing and end at nnz");
  if (h[5]) throw std::runtime_error("CSR operator: negative column index (after rebasing)");
  A.max_col = h[4];
  if (h[2]) fprintf(stderr, "[krylov_b200] warning: CSR column indices are not strictly ascending within rows; "
                            "results remain correct but are no longer bit-comparable to SparseArrays' order\n");
  A.tile_cap = h[0];
  A.max_row = h[1];
  // Ring sizing: prefer 2 CTAs/SM (<= 110 KB each) with up to 4 stages; fall
  // back to 1 CTA/SM (<= 220 KB) with >= 2 stages; otherwise no TMA staging.
  TileLayout<T> L{A.tile_cap};
  const size_t two_cta = 110 * 1024, one_cta = 220 * 1024;
  A.tma_ok = false;
  int per_sm = 2;
  // tuning overrides (profiles/sweep_k1.py): KB200_STAGES, KB200_CTAS_PER_SM
  const char* es = getenv("KB200_STAGES");
  const char* ec = getenv("KB200_CTAS_PER_SM");
  if (es && ec) {
    const int s = atoi(es), cps = atoi(ec);
    if (s >= 1 && s <= 8 && cps >= 1 && cps <= 8 && L.total_bytes(s) * cps <= 226 * 1024) { A.tma_ok = true; A.stages = s; per_sm = cps; }
  }
  // default: 3 CTAs/SM x 2 stages when it fits.  Measured on cfg2 (profiles/r1_sweep_k1.txt): K1 = 219 us at
  // 2 stages x 3 CTAs, 227 us at 3x3, 284-298 us at any depth with 2 CTAs: the gather latency wants 27 warps/SM,
  // and a shallower ring leaves more of the 228 KB to L1 for the gathered vectors.
  if (!A.tma_ok && L.total_bytes(2) * 3 <= 226 * 1024) { A.tma_ok = true; A.stages = 2; per_sm = 3; }
  for (int s = 4; s >= 2 && !A.tma_ok; s--)
    if (L.total_bytes(s) <= two_cta) { A.tma_ok = true; A.stages = s; per_sm = 2; }
  for (int s = 4; s >= 2 && !A.tma_ok; s--)
    if (L.total_bytes(s) <= one_cta) { A.tma_ok = true; A.stages = s; per_sm = 1; }
  if (A.tma_ok) {
    A.smem_bytes = L.total_bytes(A.stages);
    int g = sm_count() * per_sm;
    A.grid = g < A.ntiles ? g : (A.ntiles > 0 ? A.ntiles : 1);
    A.ctas_per_sm = per_sm;
  } else {
    A.stages = 0; A.smem_bytes = 0; A.grid = 0;
  }
}

// ---------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------
// Row-per-thread LDG kernel: always valid (any row length), used when the
// staging plan does not fit and as an independent check of the staged kernel.
template <class T, bool DOT, class G>
__global__ void __launch_bounds__(kBlock) spmv_rows_kernel(Csr<T> A, G xg, T* __restrict__ y, T* part,
                                                           unsigned* ticket, T* out) {
  __shared__ T sm[32];
  T dacc = T(0);
  const int stride = gridDim.x * blockDim.x;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < A.n; row += stride) {
    const int kb = A.rowptr[row], ke = A.rowptr[row + 1];
    T acc = T(0);
    for (int k = kb; k < ke; k++) acc = add_rn(acc, mul_rn(A.val[k], xg(A.colind[k])));
    y[row] = acc;
    if (DOT) dacc += __ldg(&xg.x[row]) * acc;
  }
  if (DOT) {
    T mine[1] = {block_sum(dacc, sm)}, tot[1];
    if (grid_sum_last<T, 1>(mine, part, ticket, sm, tot) && threadIdx.x == 0) out[0] = tot[0];
  }
}

template <class T, bool DOT, class G>
__global__ void __launch_bounds__(kTileThreads, 3) spmv_tma_kernel(Csr<T> A, G xg, T* __restrict__ y, T* part,
                                                                unsigned* ticket, T* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ T sm[32];
  T dacc = T(0);
  spmv_tiles_run<T>(
      A, smem, xg, [&](int row) { return DOT ? __ldg(&xg.x[row]) : T(0); },
      [&](int row, T acc, T xr) {
        y[row] = acc;
        if (DOT) dacc += xr * acc;
      });
  if (DOT) {
    T mine[1] = {block_sum(dacc, sm)}, tot[1];
    if (grid_sum_last<T, 1>(mine, part, ticket, sm, tot) && threadIdx.x == 0) out[0] = tot[0];
  }
}

template <class T, bool DOT, class G>
static void spmv_launch_g(Ctx& c, const Csr<T>& A, G xg, T* y, int slot, int variant) {
  T* out = reinterpret_cast<T*>(reinterpret_cast<double*>(c.dscal) + slot);
  const bool staged = variant == 2 || (variant == 0 && A.tma_ok);
  if (staged) {
    if (!A.tma_ok) throw std::runtime_error("TMA-staged SpMV requested but the tile plan does not fit shared memory");
    ensure_dyn_smem((const void*)spmv_tma_kernel<T, DOT, G>, 220 * 1024);
    int occ = 0;   // persistent grid = what is really co-resident (never more than one wave)
    KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_tma_kernel<T, DOT, G>, kTileThreads, A.smem_bytes));
    if (occ < 1) throw std::runtime_error("spmv_tma_kernel does not fit on an SM with the planned shared-memory ring");
    const int grid = std::min(std::min(occ, A.ctas_per_sm) * sm_count(), std::max(1, A.ntiles));
    spmv_tma_kernel<T, DOT, G><<<grid, kTileThreads, A.smem_bytes, c.stream>>>(A, xg, y, (T*)c.partials, c.tickets + 1, out);
  } else {
    const int grid = stream_grid(A.n, 1, 8);
    spmv_rows_kernel<T, DOT, G><<<grid, kBlock, 0, c.stream>>>(A, xg, y, (T*)c.partials, c.tickets + 1, out);
  }
  KB_CUDA(cudaGetLastError());
  c.launches++;
}

template <class T, bool DOT>
static void spmv_launch(Ctx& c, const Csr<T>& A, const T* x, T* y, int slot, int variant) {
  if (A.n <= 0) return;
  if (c.dex) spmv_launch_g<T, DOT, XGather<T>>(c, A, xgather_of<T>(c, x), y, slot, variant);   // row-partitioned: [local | halo]
  else spmv_launch_g<T, DOT, XPlain<T>>(c, A, XPlain<T>{x}, y, slot, variant);
}

template <class T> void k_spmv(Ctx& c, const Csr<T>& A, const T* x, T* y, int variant) { spmv_launch<T, false>(c, A, x, y, 1, variant); }

// ---------------------------------------------------------------------------
// General x-halo exchange of row-partitioned operators (dist.cuh: DistExchange)
// ---------------------------------------------------------------------------
template <class T> struct ExchangeDst { T* p[kMaxRanks]; };

template <class T>
__global__ void __launch_bounds__(kBlock) halo_exchange_kernel(const T* __restrict__ x, int nsend, const int* __restrict__ row,
                                                               const int* __restrict__ peer, const int* __restrict__ slot,
                                                               ExchangeDst<T> dst, unsigned* ticket, DistComm* dc) {
  __shared__ bool is_last;
  const int stride = gridDim.x * blockDim.x;
  bool sent = false;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nsend; e += stride) {
    dst.p[peer[e]][slot[e]] = x[row[e]];
    sent = true;
  }
  if (sent) __threadfence_system();        // my stores are visible to the peers before the barrier below
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  // the last CTA of this rank enters the cross-GPU barrier: when it returns, every rank has finished pushing
  if (is_last && threadIdx.x == 0) dist_allreduce_sum(dc, 0.0);
}

template <class T> void k_halo_exchange(Ctx& c, const T* x) {
  if (!c.dex) return;
  DistExchange& d = *c.dex;
  ExchangeDst<T> dst;
  const size_t par = (size_t)(d.count & 1);
  for (int k = 0; k < kMaxRanks; k++)
    dst.p[k] = d.xhalo_peer[k] ? reinterpret_cast<T*>(d.xhalo_peer[k]) + par * (size_t)d.nhalo_peer[k] : nullptr;
  const int grid = d.nsend > 0 ? std::min(sm_count(), (d.nsend + kBlock - 1) / kBlock) : 1;
  halo_exchange_kernel<T><<<grid, kBlock, 0, c.stream>>>(x, d.nsend, d.send_row, d.send_peer, d.send_slot, dst, c.tickets + 6, c.dcomm);
  KB_CUDA(cudaGetLastError());
  c.launches++;
  d.count++;
}

// ---------------------------------------------------------------------------
// Operator application (A, M, N as the solvers see them)
// ---------------------------------------------------------------------------
template <class T> void op_apply(Ctx& c, const LinOp<T>& op, const T* x, T* y, bool ldiv) {
  switch (op.kind) {
    case LinOp<T>::CSR:
      k_halo_exchange<T>(c, x);            // row-partitioned operators only; no-op on a single GPU
      k_spmv<T>(c, *op.csr, x, y, 0);
      break;
    case LinOp<T>::DIAG: k_diagmul<T>(c, op.n, y, op.diag, x, ldiv); break;
    case LinOp<T>::BDIAG: k_blockdiag_mul<T>(c, op.n, op.bs, ldiv ? op.blocks_inv : op.blocks, x, y); break;
    case LinOp<T>::DEV_CB:
      c.sync();                       // the callback may use its own stream
      op.fn(x, y, op.userdata);
      KB_CUDA(cudaDeviceSynchronize());
      break;
    case LinOp<T>::HOST_CB:
      // reference: ccall(op.fptr, ..., x, y, userdata) on host pointers
      // (interfaces/src/c_operator.jl:35-42); here x/y live in HBM, so stage.
      KB_CUDA(cudaMemcpyAsync(op.hx, x, sizeof(T) * (size_t)op.n, cudaMemcpyDeviceToHost, c.stream));
      c.sync();
      op.fn(op.hx, op.hy, op.userdata);
      KB_CUDA(cudaMemcpyAsync(y, op.hy, sizeof(T) * (size_t)op.n, cudaMemcpyHostToDevice, c.stream));
      break;
    case LinOp<T>::NONE: k_copy<T>(c, op.n, y, x); break;
  }
}

#define INST(T)                                                                                                  \
  template void csr_upload<T>(Ctx&, Csr<T>&, int, long long, const void*, const void*, const T*, int, int, bool); \
  template void csr_free<T>(Csr<T>&);                                                                            \
  template void csr_plan<T>(Ctx&, Csr<T>&);                                                                      \
  template void k_spmv<T>(Ctx&, const Csr<T>&, const T*, T*, int);                                               \
  template void k_halo_exchange<T>(Ctx&, const T*);                                                              \
  template void op_apply<T>(Ctx&, const LinOp<T>&, const T*, T*, bool);
INST(double)
INST(float)
#undef INST

}  // namespace kb
