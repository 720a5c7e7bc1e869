"""K-sharded MPPI over torch.distributed ranks (one process per GPU) -- SURVEY.md section 8e.

Rollouts are independent given (x0, U, goal), so GPU g of G rolls the global rollout ids
[offset_g, offset_g + K_g).  The Philox counter is the GLOBAL id, hence the sampled noise -- and the
result up to summation order -- does not depend on G.  Per step there is exactly one exchange: an
all-gather of the per-rank record (T x 6 float64 = 3 KB at T=64: min V, sum e, sum e*eps0, sum e*eps1,
sum eps0, sum eps1 per t).  Every rank then merges the G records (rescaling by exp(-(m_g-m)/lam)) and
finishes clip / SavGol / clip / shift redundantly, so all ranks hold the identical new U and no
broadcast is needed.  (A plain sum-allreduce as BASELINE.json words it is only correct for a
pre-agreed shift; with lam=1e-3 a fixed shift over/underflows, so the min travels in the record.)
"""
import ctypes as C

import numpy as np

from . import _capi
from .mppi import MPPI


def shard_plan(K_total, world, rank):
    """(K_local, k_offset): contiguous split, the first K_total % world ranks take one extra rollout."""
    base, rem = divmod(int(K_total), int(world))
    k_local = base + (1 if rank < rem else 0)
    k_offset = rank * base + min(rank, rem)
    return k_local, k_offset


def merge_records(records, lam, eps_floor, K_total):
    """Reference (NumPy) statement of the merge finalize_kernel performs: records (G,T,6) -> dU (2,T)."""
    records = np.asarray(records, dtype=np.float64)
    m = records[:, :, 0].min(axis=0)
    sc = np.exp(-(records[:, :, 0] - m[None, :]) / lam)
    S = (records[:, :, 1] * sc).sum(axis=0)
    N = (records[:, :, 2:4] * sc[:, :, None]).sum(axis=0)
    E = records[:, :, 4:6].sum(axis=0)
    return ((N + eps_floor * E) / (S[:, None] + eps_floor * K_total)).T


class _DevBuf(object):
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n_f64):
        self.__cuda_array_interface__ = {"shape": (n_f64,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class ShardedMPPI(object):
    """`MPPI` whose `samples` are sharded over the ranks of a torch.distributed process group.

    exchange='p2p'  (default with the nccl backend): the ranks map each other's row buffers through CUDA IPC once;
                    afterwards every reduce block stores the row of its time step straight into every peer's memory over
                    NVLink as flag-in-data words, merges the peers' rows of the same step, and the finalizer block of each
                    rank's reduce kernel finishes the update -- two kernels per rank and no collective call per step.
    exchange='nccl' all-gathers the device-resident records in place with NCCL (engine kernels and the
                    collective share one stream, no host sync in between).
    exchange='host' stages the 3 KB record through host memory (works with any backend, e.g. gloo).
    """

    def __init__(self, horizon, samples_total, group=None, exchange=None, **engine):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.samples_total = int(samples_total)
        k_local, k_offset = shard_plan(samples_total, self.world, self.rank)
        backend = dist.get_backend(group)
        self.exchange = exchange or ("p2p" if backend == "nccl" and self.world > 1 else "host")
        self.horizon = horizon
        self._engine_opts = dict(engine)
        self._create_engine(k_local, k_offset)
        if self.exchange == "p2p":
            # every rank must end up on the same transport: agree on whether the peer mappings succeeded
            ok = 1
            try:
                self._connect_p2p(backend)
            except (_capi.MppiError, RuntimeError) as ex:
                ok, self.p2p_error = 0, str(ex)
            flag = torch.tensor([ok], dtype=torch.int32)
            if backend == "nccl":
                flag = flag.cuda()
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                # no peer access between some pair of GPUs (or IPC unavailable): fall back to the collective
                self.mppi.close()
                self.exchange = "nccl" if backend == "nccl" else "host"
                self._create_engine(k_local, k_offset)
            dist.barrier(group=group)
        if self.exchange == "nccl":
            self._rec_t = torch.as_tensor(_DevBuf(self._rec_ptr, self._n_rec), device="cuda")
            self._gat_t = torch.as_tensor(_DevBuf(self._gat_ptr, self._n_rec * self.world), device="cuda")

    def _create_engine(self, k_local, k_offset):
        torch = self._torch
        engine = dict(self._engine_opts)
        self.stream = None
        if self.exchange == "nccl":
            # a dedicated non-default stream: the engine launches its kernels on it and the NCCL collective
            # is enqueued under the same stream context, so kernels and all-gather are stream-ordered
            # (handle 0 = the legacy default stream cannot be passed through the C ABI: NULL means
            # "engine-owned stream")
            self.stream = torch.cuda.Stream()
            engine.setdefault("stream", self.stream.cuda_stream)
        self.mppi = MPPI(horizon=self.horizon, samples=k_local, k_offset=k_offset, k_total=self.samples_total,
                         world_size=self.world, rank=self.rank, **engine)
        lib, h = self.mppi._lib, self.mppi._h
        rec, gat = C.c_void_p(), C.c_void_p()
        rb, gb = C.c_size_t(), C.c_size_t()
        _capi.check(lib.mppi_exchange_buffers(h, C.byref(rec), C.byref(rb), C.byref(gat), C.byref(gb)), "mppi_exchange_buffers")
        self._n_rec = rb.value // 8
        self._rec_ptr, self._gat_ptr = rec.value, gat.value

    def _connect_p2p(self, backend):
        """Map every peer's exchange buffer (CUDA IPC) so that the reduce kernel can store its record there."""
        torch, dist, group = self._torch, self._dist, self.group
        lib, h = self.mppi._lib, self.mppi._h
        mine = (C.c_ubyte * 64)()
        st = lib.mppi_p2p_export(h, C.cast(mine, C.c_void_p))
        t = torch.tensor(list(bytes(mine)) + [0 if st == 0 else 1], dtype=torch.uint8)
        if backend == "nccl":
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=group)         # every rank takes part, also one whose export failed
        allb = torch.stack([o.cpu() for o in out])
        if int(allb[:, 64].max()) != 0:
            raise RuntimeError("mppi_p2p_export failed on a rank")
        allh = bytes(allb[:, :64].reshape(-1).tolist())
        buf = (C.c_ubyte * len(allh)).from_buffer_copy(allh)
        _capi.check(lib.mppi_p2p_connect(h, C.cast(buf, C.c_void_p)), "mppi_p2p_connect")

    def get_path(self, state, goal, sig=np.array([[.9, 0.0], [0.0, .9]]), lam=.001):
        """MPPI.get_path (control/src/mppi:85-102) with K sharded over the group."""
        m, lib, h = self.mppi, self.mppi._lib, self.mppi._h
        m._sync_sampling(sig, lam)
        if self.exchange == "p2p":
            return m.get_path(state, goal, sig, lam)      # plain mppi_step: the exchange is inside the graph
        _capi.check(lib.mppi_set_goal(h, _capi.dptr(_capi.f64(goal, (3,)))), "mppi_set_goal")
        m._goal_bytes = None
        u, x = np.empty(2), np.empty(3)
        for attempt in range(3):
            _capi.check(lib.mppi_step_local(h, _capi.dptr(_capi.f64(state, (3,)))), "mppi_step_local")
            if self.exchange == "nccl":
                with self._torch.cuda.stream(self.stream):
                    self._dist.all_gather_into_tensor(self._gat_t, self._rec_t, group=self.group)
            else:
                rec = np.empty(self._n_rec)
                _capi.check(lib.mppi_read_record(h, _capi.dptr(rec)), "mppi_read_record")
                allrec = exchange_host(self._torch, self._dist, rec, self.world, self.group)
                _capi.check(lib.mppi_write_gather(h, _capi.dptr(allrec)), "mppi_write_gather")
            st = lib.mppi_step_finish(h, _capi.dptr(u), _capi.dptr(x))
            # MPPI_ERR_RETRY: the fp32 screen of precision 'mixed' overflowed on some rank (systematic within centimetres of
            # the goal).  Every rank merged the same records, so every rank is here for the same step: nothing was applied,
            # the second round runs the fp64 pipeline on the same noise.
            if st != _capi.MPPI_ERR_RETRY:
                break
        _capi.check(st, "mppi_step_finish")
        m._path_log.append(x)
        m._uvec_log.append(u)
        m.fin_time.append(m.fin_time[-1] + m.dt)
        return x

    def initialize(self):
        self.mppi.initialize()

    @property
    def uvec(self):
        return self.mppi.uvec

    @property
    def latest_uvec(self):
        return self.mppi.latest_uvec


def exchange_host(torch, dist, rec, world, group=None):
    """all-gather of one host record per rank -> (world * n) float64, rank-major."""
    t = torch.from_numpy(np.ascontiguousarray(rec, dtype=np.float64))
    if dist.get_backend(group) == "nccl":          # NCCL moves device tensors only
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.concatenate([o.cpu().numpy() for o in out])
