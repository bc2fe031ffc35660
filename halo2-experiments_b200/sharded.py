"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed
for the plumbing (NCCL on GPUs, gloo in the CPU tests).

* `ShardedCommit`  — one best_multiexp / ParamsKZG::commit split by point range: rank g holds
  bases[lo_g, hi_g) resident on its GPU, multiplies its slice of the scalars, and the G partial
  G1 points (96 bytes each) are all-gathered and added.  EC addition is not an NCCL reduce op,
  hence gather-then-add; the exchange is latency-bound (G x 96 B).
* `FourStepNTT`    — one best_fft of size N = R*C split over G ranks: rank g owns a block of C/G
  columns, does the R-point column transforms and the inter-step twiddles locally, an all-to-all
  transposes blocks so that rank g owns R/G full rows, then does the C-point row transforms.
  Output element k = k_r + R*k_c ends up on the rank that owns row k_r (digit-reversed
  distribution, which is what a row-sharded quotient evaluation consumes).

The local arithmetic is injected (`engine`), so the same host logic runs on the GPU backend and,
in tests/test_sharded_cpu.py, on a CPU engine under gloo with world size 2.
"""
import numpy as np


def shard_range(n, world, rank):
    """Contiguous point range [lo, hi) of rank `rank` (sizes differ by at most one)."""
    return n * rank // world, n * (rank + 1) // world


def _torch():
    import torch
    import torch.distributed as dist
    return torch, dist


def all_gather_u64(arr, device=None):
    """All-gather equal-sized uint64 arrays; returns (world, *arr.shape).  Works for gloo (CPU
    tensors) and NCCL (device = torch.device('cuda', i))."""
    torch, dist = _torch()
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return a[None]
    t = torch.from_numpy(a.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in outs])


def all_to_all_u64(blocks, device=None):
    """blocks[j] goes to rank j; returns the list of blocks received (one per rank)."""
    torch, dist = _torch()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [np.ascontiguousarray(blocks[0], dtype=np.uint64)]
    send = [torch.from_numpy(np.ascontiguousarray(b, dtype=np.uint64).view(np.int64).copy()) for b in blocks]
    if device is not None:
        send = [s.to(device) for s in send]
    recv = [torch.empty_like(s) for s in send]
    if dist.get_backend() == "gloo":                     # gloo has no all_to_all: pairwise exchange
        rank, world = dist.get_rank(), dist.get_world_size()
        recv[rank] = send[rank].clone()
        for step in range(1, world):
            dst, src = (rank + step) % world, (rank - step) % world
            req = dist.isend(send[dst], dst)
            dist.recv(recv[src], src)
            req.wait()
    else:
        dist.all_to_all(recv, send)
    return [r.cpu().numpy().view(np.uint64) for r in recv]


class ShardedCommit:
    """Point-range sharded MSM.  engine.msm(scalars) -> (12,) uint64 partial G1 over this rank's bases;
    engine.g1_sum(points (G,12)) -> (12,) uint64."""

    def __init__(self, engine, device=None):
        self.engine, self.device = engine, device

    def commit(self, local_scalars):
        partial = self.engine.msm(local_scalars)
        gathered = all_gather_u64(partial, self.device)
        return self.engine.g1_sum(gathered)


class GpuCommitEngine:
    """ShardedCommit engine over a device-resident SRS slice (ParamsKZG of the local range)."""

    def __init__(self, zk, params, lagrange=False):
        self.zk, self.params, self.lagrange = zk, params, lagrange

    def msm(self, scalars):
        if hasattr(scalars, "ptr"):                       # DeviceBuffer
            return self.params.commit_dev(scalars, self.params.n, self.lagrange)
        return self.params.commit_lagrange(scalars) if self.lagrange else self.params.commit(scalars)

    def g1_sum(self, points):
        import ctypes
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
        out = np.zeros(12, dtype=np.uint64)
        rc = self.zk.lib().b200zk_g1_sum(pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(pts.shape[0]),
                                         out.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise self.zk.B200zkError(f"b200zk_g1_sum failed: {rc}")
        return out


class GpuNttEngine:
    """FourStepNTT engine over one Backend with host (numpy) blocks: upload, kernel, download.
    Used to check the kernels with the ranks emulated one after another on a single GPU."""

    def __init__(self, zk, backend, log_n):
        self.zk, self.be, self.log_n = zk, backend, log_n

    def col_step(self, block, omega_n, log_r, log_c, col0):
        import ctypes
        block = np.ascontiguousarray(block, dtype=np.uint64)
        log_cg = int(np.log2(block.shape[1]))
        d = self.be.to_device(block)
        try:
            self.be._check(self.zk.lib().b200zk_fft_colstep_dev(self.be._ctx, d.ptr, ctypes.c_uint32(log_r), ctypes.c_uint32(log_cg),
                                                              ctypes.c_uint32(col0), np.ascontiguousarray(omega_n).ctypes.data_as(ctypes.c_void_p),
                                                              ctypes.c_uint32(self.log_n)))
            return d.download(block.shape)
        finally:
            d.free()

    def row_step(self, rows, omega_c, log_c):
        import ctypes
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        d = self.be.to_device(rows)
        try:
            self.be._check(self.zk.lib().b200zk_fft_rows_dev(self.be._ctx, d.ptr, ctypes.c_uint32(rows.shape[0]),
                                                           np.ascontiguousarray(omega_c).ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(log_c)))
            return d.download(rows.shape)
        finally:
            d.free()


class FourStepNTTDevice:
    """The same four-step transform with the data resident in HBM: blocks are torch CUDA tensors
    (int64 views of the u64 limbs, torch owns the memory), the kernels run through the C ABI on the
    tensors' device pointers and the transpose is one NCCL all_to_all_single over NVLink."""

    def __init__(self, zk, backend, log_n, log_r, rank, world, device):
        import ctypes
        self.ct = ctypes
        self.zk, self.be, self.log_n, self.log_r, self.log_c = zk, backend, log_n, log_r, log_n - log_r
        self.rank, self.world, self.device = rank, world, device
        self.R, self.C = 1 << log_r, 1 << (log_n - log_r)
        assert self.R % world == 0 and self.C % world == 0

    def forward(self, block, omega_n, omega_c):
        """block: torch int64 CUDA tensor (R, C/G, 4), overwritten.  Returns (R/G, C, 4)."""
        torch, dist = _torch()
        ct, lib = self.ct, self.zk.lib()
        cg, rg = self.C // self.world, self.R // self.world
        torch.cuda.synchronize(self.device)
        self.be._check(lib.b200zk_fft_colstep_dev(self.be._ctx, ct.c_void_p(block.data_ptr()), ct.c_uint32(self.log_r),
                                                 ct.c_uint32(int(np.log2(cg))), ct.c_uint32(self.rank * cg),
                                                 np.ascontiguousarray(omega_n).ctypes.data_as(ct.c_void_p), ct.c_uint32(self.log_n)))
        self.be.sync()
        if self.world > 1:
            recv = torch.empty_like(block)                       # (G, rg, cg, 4) chunks by source rank
            dist.all_to_all_single(recv, block)
            rows = recv.view(self.world, rg, cg, 4).permute(1, 0, 2, 3).contiguous().view(rg, self.C, 4)
        else:
            rows = block.view(rg, self.C, 4)
        torch.cuda.synchronize(self.device)
        self.be._check(lib.b200zk_fft_rows_dev(self.be._ctx, ct.c_void_p(rows.data_ptr()), ct.c_uint32(rg),
                                              np.ascontiguousarray(omega_c).ctypes.data_as(ct.c_void_p), ct.c_uint32(self.log_c)))
        self.be.sync()
        return rows


class FourStepNTTFused(FourStepNTTDevice):
    """Same transform with the exchange fused into the column-step kernel: every rank exports its
    (R/G, C) row buffer over CUDA IPC, the column step writes each transformed row straight into the
    owner's buffer with NVLink / NVSwitch peer stores (b200zk_fft_colstep_scatter_dev), and after a
    barrier every rank runs the row step on what it received.  No NCCL all-to-all, no staging copy,
    no transpose pass."""

    def __init__(self, zk, backend, log_n, log_r, rank, world, device):
        super().__init__(zk, backend, log_n, log_r, rank, world, device)
        torch, dist = _torch()
        ct, lib = self.ct, zk.lib()
        rg = self.R // world
        self.rows_buf = backend.alloc(rg * self.C * 32)
        handle = (ct.c_uint8 * 64)()
        backend._check(lib.b200zk_ipc_get_handle(backend._ctx, self.rows_buf.ptr, handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        self.peers = (ct.c_void_p * world)()
        self._opened = []
        for j in range(world):
            if j == rank:
                self.peers[j] = self.rows_buf.ptr.value
            else:
                p = ct.c_void_p()
                backend._check(lib.b200zk_ipc_open(backend._ctx, (ct.c_uint8 * 64).from_buffer_copy(handles[j]), ct.byref(p)))
                self.peers[j] = p.value
                self._opened.append(p)
        dist.barrier()

    def forward(self, block, omega_n, omega_c):
        """block: torch int64 CUDA tensor (R, C/G, 4), read only.  Returns this rank's DeviceBuffer with
        the (R/G, C, 4) output rows."""
        torch, dist = _torch()
        ct, lib = self.ct, self.zk.lib()
        cg, rg = self.C // self.world, self.R // self.world
        torch.cuda.synchronize(self.device)
        self.be._check(lib.b200zk_fft_colstep_scatter_dev(self.be._ctx, ct.c_void_p(block.data_ptr()), ct.c_uint32(self.log_r),
                                                         ct.c_uint32(int(np.log2(cg))), ct.c_uint32(self.rank * cg),
                                                         np.ascontiguousarray(omega_n).ctypes.data_as(ct.c_void_p), ct.c_uint32(self.log_n),
                                                         self.peers, ct.c_uint32(self.world)))
        self.be.sync()
        dist.barrier()                                          # every rank's rows have arrived
        self.be._check(lib.b200zk_fft_rows_dev(self.be._ctx, self.rows_buf.ptr, ct.c_uint32(rg),
                                              np.ascontiguousarray(omega_c).ctypes.data_as(ct.c_void_p), ct.c_uint32(self.log_c)))
        self.be.sync()
        dist.barrier()                                          # nobody may overwrite a peer's rows before it has transformed them
        return self.rows_buf

    def close(self):
        lib = self.zk.lib()
        for p in self._opened:
            lib.b200zk_ipc_close(self.be._ctx, p)
        self._opened = []
        self.rows_buf.free()


class FourStepNTT:
    """Row/column sharded four-step NTT of size N = R*C over `world` ranks (R, C powers of two,
    world divides both).

    engine.col_step(block, omega_n, log_r, log_c, col0) : block is (R, C/G, 4): R-point transforms down
        the columns followed by the twiddle omega_n^(c_global * k_r); returns the same shape.
    engine.row_step(rows, omega_c, log_c)               : rows is (R/G, C, 4): C-point transform of each row.
    """

    def __init__(self, engine, log_n, log_r, rank, world, device=None):
        self.engine, self.log_n, self.log_r, self.log_c = engine, log_n, log_r, log_n - log_r
        self.rank, self.world, self.device = rank, world, device
        self.R, self.C = 1 << log_r, 1 << (log_n - log_r)
        assert self.R % world == 0 and self.C % world == 0

    def local_columns(self, full):
        """This rank's input block (R, C/G, 4) cut out of a natural-order array (tests / host callers)."""
        cg = self.C // self.world
        return np.ascontiguousarray(np.asarray(full).reshape(self.R, self.C, 4)[:, self.rank * cg:(self.rank + 1) * cg])

    def forward(self, block, omega_n, omega_c):
        """block: (R, C/G, 4) columns owned by this rank.  Returns (R/G, C, 4): rows k_r in this rank's
        row range, entry [k_r_local, k_c] = X[k_r + R * k_c]."""
        cg, rg = self.C // self.world, self.R // self.world
        y = self.engine.col_step(block, omega_n, self.log_r, self.log_c, self.rank * cg)
        send = [np.ascontiguousarray(y[j * rg:(j + 1) * rg]) for j in range(self.world)]      # (R/G, C/G, 4) per peer
        recv = all_to_all_u64(send, self.device)
        rows = np.concatenate([r.reshape(rg, cg, 4) for r in recv], axis=1)                   # (R/G, C, 4)
        return self.engine.row_step(rows, omega_c, self.log_c)

    def gather_natural(self, rows):
        """All-gather the row blocks and restore natural order (tests only)."""
        allr = all_gather_u64(rows, self.device).reshape(self.R, self.C, 4)                    # [k_r][k_c]
        return np.ascontiguousarray(np.transpose(allr, (1, 0, 2))).reshape(self.R * self.C, 4) # k = k_r + R*k_c
