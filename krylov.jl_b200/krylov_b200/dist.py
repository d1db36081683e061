"""Row-partitioned (multi-GPU) CG: one process per GPU under torchrun.

torch.distributed is plumbing only (rendezvous, exchanging 256-byte CUDA-IPC
handle blobs, the max-over-ranks of the timing); the data path is inside the
CUDA kernels: halo entries of r and p are loaded from the peers' HBM over
NVLink while K1 runs, and the two dot products per iteration finish with an
in-kernel all-reduce through peer mailboxes (csrc/dist.cuh).

The reference has no distributed code; docs/src/custom_workspaces.md:464-637
sketches the same decomposition with MPI (local dot + Allreduce, user-written
distributed mul!).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib


# --------------------------------------------------------------------------
# partitioning (pure NumPy: testable on CPU with gloo)
# --------------------------------------------------------------------------
def slab_bounds(n3: int, world: int):
    """Balanced split of n3 z-planes over `world` ranks -> list of (k_lo, k_hi)."""
    base, rem = divmod(n3, world)
    out, k = [], 0
    for r in range(world):
        h = base + (1 if r < rem else 0)
        out.append((k, k + h))
        k += h
    return out


def localize_columns(colind_global, row_starts, rank):
    """Map GLOBAL column indices of a row block to [local | halo] numbering.

    row_starts: array of length world+1 with the first global row of every rank.
    Returns (colind_local int32, halo_rank int32[nhalo], halo_off int32[nhalo]).
    Column j owned by this rank -> j - lo; any other column -> nloc + h where h indexes the sorted
    unique list of off-rank columns this block touches ("bit-exact integer indexing": the map is a
    bijection on the touched columns and preserves the ascending order inside every row)."""
    xp_is_torch = type(colind_global).__module__.startswith("torch")
    if xp_is_torch:
        import torch
        rs = torch.as_tensor(np.asarray(row_starts), device=colind_global.device, dtype=torch.int64)
        lo, hi = int(row_starts[rank]), int(row_starts[rank + 1])
        cg = colind_global.to(torch.int64)
        off = (cg < lo) | (cg >= hi)
        halo_cols = torch.unique(cg[off])                      # sorted
        owner = torch.searchsorted(rs, halo_cols, right=True) - 1
        halo_off = halo_cols - rs[owner]
        local = cg - lo
        if halo_cols.numel():
            local[off] = (hi - lo) + torch.searchsorted(halo_cols, cg[off])
        return (local.to(torch.int32), owner.to(torch.int32).cpu().numpy(), halo_off.to(torch.int32).cpu().numpy())
    rs = np.asarray(row_starts, dtype=np.int64)
    lo, hi = int(rs[rank]), int(rs[rank + 1])
    cg = np.asarray(colind_global, dtype=np.int64)
    off = (cg < lo) | (cg >= hi)
    halo_cols = np.unique(cg[off])
    owner = np.searchsorted(rs, halo_cols, side="right") - 1
    halo_off = halo_cols - rs[owner]
    local = cg - lo
    if len(halo_cols):
        local[off] = (hi - lo) + np.searchsorted(halo_cols, cg[off])
    return local.astype(np.int32), owner.astype(np.int32), halo_off.astype(np.int32)


def push_ranges(maps, rank, max_ranges=4):
    """Send plan of `rank` from every rank's halo map (halo_rank, halo_off).

    For each peer q, the halo slots of q that `rank` owns must form ONE run of consecutive slots pointing at
    consecutive local rows; returns [(first local row, count, q, first slot in q's halo)] or None when some
    peer's needs are not contiguous (the kernels then stay in pull mode)."""
    out = []
    for q, (hr, ho) in enumerate(maps):
        if q == rank:
            continue
        hr, ho = np.asarray(hr), np.asarray(ho)
        slots = np.nonzero(hr == rank)[0]
        if len(slots) == 0:
            continue
        rows = ho[slots]
        if not (np.all(np.diff(slots) == 1) and np.all(np.diff(rows) == 1)):
            return None
        out.append((int(rows[0]), int(len(slots)), int(q), int(slots[0])))
    if len(out) > max_ranges:
        return None
    return out


def send_list(maps, rank):
    """General send list of `rank` from every rank's halo map: for each peer q and each halo slot h of q owned by
    `rank`, one entry (local row = halo_off_q[h], peer = q, slot = h).  Returns three int32 arrays."""
    rows, peers, slots = [], [], []
    for q, (hr, ho) in enumerate(maps):
        if q == rank:
            continue
        hr, ho = np.asarray(hr), np.asarray(ho)
        idx = np.nonzero(hr == rank)[0]
        rows.append(ho[idx]); peers.append(np.full(len(idx), q)); slots.append(idx)
    cat = lambda a: np.ascontiguousarray(np.concatenate(a) if a else np.zeros(0), dtype=np.int32)
    return cat(rows), cat(peers), cat(slots)


# --------------------------------------------------------------------------
# distributed workspace
# --------------------------------------------------------------------------
class DistWorkspace:
    """Workspace of any of the four solvers for one row block.  `csr_local` = (rowptr, colind_local, values) with
    the [local | halo] column numbering of localize_columns().  CG runs its fused two-kernel iteration with the
    in-kernel halo pull; every other product (all solvers, primitive and fused-phase paths) is preceded by the
    general halo exchange kernel; every dot product ends in the in-kernel all-reduce."""

    def __init__(self, solver, csr_local, halo_rank, halo_off, rank, world, dtype=np.float64, device="cuda",
                 memory=0, window=0):
        import torch.distributed as dist
        from . import krylov_workspace
        self.rank, self.world = rank, world
        nloc = int(csr_local[0].shape[0]) - 1
        self.ws = krylov_workspace(solver, nloc, nloc, dtype, device=device, memory=memory, window=window)
        self.ws.set_operator(csr_local)
        L = _lib.lib()
        hr = np.ascontiguousarray(halo_rank, dtype=np.int32)
        ho = np.ascontiguousarray(halo_off, dtype=np.int32)
        if L.krylov_b200_dist_init(self.ws._h, rank, world, len(hr), hr.ctypes.data_as(C.c_void_p),
                                   ho.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError(_lib.last_error())
        # push mode: if what every peer needs from me is a few contiguous row ranges, producers store those
        # entries straight into the peers' halo buffers (no fine-grained P2P loads in the SpMV)
        maps = [None] * world
        dist.all_gather_object(maps, (hr, ho))
        nh = np.ascontiguousarray(np.array([len(m[0]) for m in maps], dtype=np.int32))
        nloc_all = [None] * world
        dist.all_gather_object(nloc_all, nloc)
        rows, peers, slots = send_list(maps, rank)
        if L.krylov_b200_dist_set_sendlist(self.ws._h, len(rows), rows.ctypes.data_as(C.c_void_p),
                                           peers.ctypes.data_as(C.c_void_p), slots.ctypes.data_as(C.c_void_p),
                                           nh.ctypes.data_as(C.c_void_p), int(sum(nloc_all))) != 0:
            raise RuntimeError(_lib.last_error())
        ranges = push_ranges(maps, rank) if solver == "cg" else None
        # measured at 2 GPUs (profiles/r1_scale_2gpu_push_vs_pull.txt): pull 4833 / 567.5 it/s vs push 4685 / 563.1
        # (n = 1e7 / 1e8) -- pull stays the default, push is opt-in
        self.push_mode = ranges is not None and os.environ.get("KB200_DIST_PUSH", "0") == "1"
        if self.push_mode:
            flat = np.ascontiguousarray(np.array(ranges, dtype=np.int32).reshape(-1))
            if L.krylov_b200_dist_set_push(self.ws._h, len(ranges), flat.ctypes.data_as(C.c_void_p),
                                           nh.ctypes.data_as(C.c_void_p)) != 0:
                raise RuntimeError(_lib.last_error())
        nb = L.krylov_b200_dist_handle_bytes()
        mine = (C.c_ubyte * nb)()
        if L.krylov_b200_dist_export(self.ws._h, mine) != 0:
            raise RuntimeError(_lib.last_error())
        blobs = [None] * world
        dist.all_gather_object(blobs, bytes(mine))
        allb = b"".join(blobs)
        buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
        if L.krylov_b200_dist_import(self.ws._h, buf) != 0:
            raise RuntimeError(_lib.last_error())
        dist.barrier()

    def solve(self, b_local, **kw):
        return self.ws.solve(None, b_local, **kw)

    def warm_start(self, x0_local):
        return self.ws.warm_start(x0_local)

    @property
    def x(self):
        return self.ws.x

    @property
    def stats(self):
        return self.ws.stats

    def free(self):
        self.ws.free()


class DistCgWorkspace(DistWorkspace):
    def __init__(self, csr_local, halo_rank, halo_off, rank, world, dtype=np.float64, device="cuda"):
        super().__init__("cg", csr_local, halo_rank, halo_off, rank, world, dtype=dtype, device=device)


def make_stencil_rank(gen, N, rank, world, torch, device, dtype=np.float64):
    """This rank's z-slab of a 7-point stencil matrix `gen` (problems.div_grad_csr / kron_unsymmetric_csr)."""
    bounds = slab_bounds(N, world)
    k_lo, k_hi = bounds[rank]
    rp, ci, va = gen(N, dtype=dtype, k_lo=k_lo, k_hi=k_hi, xp=torch, device=device)
    row_starts = np.array([b[0] * N * N for b in bounds] + [N ** 3], dtype=np.int64)
    ci_loc, hr, ho = localize_columns(ci, row_starts, rank)
    return (rp, ci_loc, va), hr, ho, (k_hi - k_lo) * N * N


def make_poisson_rank(N, rank, world, torch, device, dtype=np.float64):
    """This rank's z-slab of get_div_grad(N,N,N) with localized columns."""
    from .problems import div_grad_csr
    bounds = slab_bounds(N, world)
    k_lo, k_hi = bounds[rank]
    rp, ci, va = div_grad_csr(N, dtype=dtype, k_lo=k_lo, k_hi=k_hi, xp=torch, device=device)
    row_starts = np.array([b[0] * N * N for b in bounds] + [N ** 3], dtype=np.int64)
    ci_loc, hr, ho = localize_columns(ci, row_starts, rank)
    return (rp, ci_loc, va), hr, ho, (k_hi - k_lo) * N * N


# --------------------------------------------------------------------------
# bench entry (called by bench.py when WORLD_SIZE > 1)
# --------------------------------------------------------------------------
def _run_workload(N, iters, steps, warmup, rank, world, local, dev, torch, dist, sampler=None):
    """Row-partitioned fused CG on get_div_grad(N,N,N): device-resident timing (CUDA events on the workspace
    stream, max over ranks), end-to-end timing with host buffers, and the residual history of one extra solve."""
    csr, hr, ho, nloc = make_poisson_rank(N, rank, world, torch, dev)
    dws = DistCgWorkspace(csr, hr, ho, rank, world)
    b = torch.ones(nloc, dtype=torch.float64, device=dev)
    kw = dict(atol=0.0, rtol=0.0, itmax=iters)
    stream = torch.cuda.ExternalStream(_lib.lib().krylov_b200_stream(dws.ws._h), device=dev)
    torch.cuda.synchronize()
    for _ in range(warmup):
        dist.barrier()
        dws.solve(b, **kw)
    assert dws.stats.niter == iters, dws.stats
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = dws.ws.launches
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        dws.solve(b, **kw)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = torch.tensor([dws.ws.launches - l0], device=dev)
    dist.all_reduce(launches)
    # end-to-end arm: host buffers for this rank's slice of b and x
    bh = torch.ones(nloc, dtype=torch.float64).pin_memory()
    xh = torch.empty(nloc, dtype=torch.float64).pin_memory()
    bd = torch.empty(nloc, dtype=torch.float64, device=dev)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        bd.copy_(bh, non_blocking=True)                         # H2D of this rank's slice of b
        torch.cuda.synchronize()
        dws.solve(bd, **kw)
        xh.copy_(dws.x, non_blocking=False)                     # D2H of this rank's slice of x
    torch.cuda.synchronize()
    dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler is not None else None
    dws.solve(b, time_kernels=True, **kw)       # phase durations measured inside the persistent kernel (rank-local)
    k1_ms, k2_ms, timed = dws.ws.kernel_times
    # parity material (not timed): the residual history of one more solve, identical on every rank because every
    # rank derives alpha / beta / rNorm from the same all-reduced scalars -- checked here, then compared by rank 0
    dws.solve(b, history=True, **kw)
    hist = np.asarray(dws.stats.residuals, dtype=np.float64)
    hmax = torch.tensor(hist, device=dev)
    hmin = hmax.clone()
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
    ranks_agree = bool(torch.equal(hmax, hmin))
    status = dws.stats.status
    dws.free()
    del csr, b, bd
    torch.cuda.empty_cache()
    return dict(ms=ms, launches=int(launches.item()), e2e_s=float(e2e_s.item()), clocks=clocks, hist=hist,
                ranks_agree=ranks_agree, status=status,
                phases=dict(phase_a_ms=k1_ms, phase_b_ms=k2_ms, timed_iterations=timed,
                            note="rank 0, measured inside cg_persist (%globaltimer), grid barrier + cross-GPU all-reduce included"))


def bench_main(args, WORKLOADS, algorithmic_bytes_cg, hbm_peak, ClockSampler, parity_block=None, golden_parity=None, workload_name=None):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.lib().krylov_b200_set_device(local)
    dist.init_process_group("nccl", device_id=dev)
    N, iters = WORKLOADS[args.workload]
    n, nnz = N ** 3, 7 * N ** 3 - 6 * N ** 2
    sampler = ClockSampler(local) if rank == 0 else None
    res = _run_workload(N, iters, args.steps, args.warmup, rank, world, local, dev, torch, dist, sampler)
    # BASELINE config 5 (n ~ 1e8) rides along on the default workload so that the driver's 1/2/4/8-GPU scaling file
    # carries it: `cfg5.value` at 8 GPUs over `cfg5.value` at 1 GPU is north_star's ">= 6x" figure.
    cfg5 = None
    if args.workload == "poisson215" and not getattr(args, "no_cfg5", False):
        N5, it5 = WORKLOADS["poisson464"]
        r5 = _run_workload(N5, it5, max(2, args.steps // 2), 3, rank, world, local, dev, torch, dist)
        cfg5 = (N5, it5, max(2, args.steps // 2), r5)
    if rank == 0:
        its = args.steps * iters
        ms = res["ms"]
        value = its / (ms * 1e-3)
        B = algorithmic_bytes_cg(n, nnz)
        peak, src = hbm_peak()
        achieved = B * its / (ms * 1e-3) / 1e9
        parity = None
        if parity_block is not None and not getattr(args, "no_cpu", False):
            try:
                parity = golden_parity(res["hist"], "bench_cg_poisson464") if N > 300 else parity_block(res["hist"], N, iters)
            except Exception as ex:
                parity = dict(ok=None, note=f"failed: {type(ex).__name__}: {ex}")
            parity["ranks_agree"] = res["ranks_agree"]
            if not res["ranks_agree"]:
                parity["ok"] = False
        line = dict(metric="CG iterations/s", value=value, unit="it/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=workload_name(N, iters) if workload_name else f"cg! on get_div_grad({N},{N},{N})",
                                implementation=f"cg! fused: persistent cooperative kernel, int32 CSR row-partitioned in z-slabs over {world} GPUs",
                                n=n, nnz=nnz, iters_per_step=iters,
                                parallelism=f"rows/{world}: halo staged over NVLink P2P inside the kernel + in-kernel all-reduce",
                                l2="per-rank matrix slab %.0f MB" % (nnz * 12 / world / 1e6), status=res["status"]),
                    roofline=dict(bound="hbm", achieved=achieved, peak=peak * world, unit="GB/s", frac=achieved / (peak * world),
                                  traffic=None, peak_source=src + f" x {world} GPUs", bytes_per_iteration=B,
                                  kernels=res["phases"]),
                    clocks=res["clocks"],
                    e2e=dict(value=its / res["e2e_s"], unit="it/s", h2d_bytes_per_step=n * 8, d2h_bytes_per_step=n * 8),
                    gpu_launches=res["launches"])
        if parity is not None:
            line["parity"] = parity
        if cfg5 is not None:
            N5, it5, st5, r5 = cfg5
            n5, nnz5 = N5 ** 3, 7 * N5 ** 3 - 6 * N5 ** 2
            v5 = st5 * it5 / (r5["ms"] * 1e-3)
            B5 = algorithmic_bytes_cg(n5, nnz5)
            p5 = None
            if golden_parity is not None:
                p5 = golden_parity(r5["hist"], "bench_cg_poisson464")
                p5["ranks_agree"] = r5["ranks_agree"]
                if not r5["ranks_agree"]:
                    p5["ok"] = False
            line["cfg5"] = dict(workload=f"cg! on get_div_grad({N5},{N5},{N5}) (n = {n5}, nnz = {nnz5}) row-partitioned over {world} GPUs, "
                                         f"{it5} iterations per step, {st5} steps", value=v5, unit="it/s", n_gpus=world,
                                ms_per_step=r5["ms"] / st5, frac=B5 * v5 / 1e9 / (peak * world), bytes_per_iteration=B5,
                                e2e=dict(value=st5 * it5 / r5["e2e_s"], unit="it/s"), parity=p5, kernels=r5["phases"],
                                speedup_note="north_star target: value at 8 GPUs >= 6 x value at 1 GPU (same key on the N=1 line)")
        print(json.dumps(line))
        bad = [k for k, p in (("parity", line.get("parity")), ("cfg5.parity", (line.get("cfg5") or {}).get("parity"))) if p and p.get("ok") is False]
    else:
        bad = []
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        raise SystemExit("parity FAILED: " + ", ".join(bad))
