"""CPU, world_size 2, gloo: the N>1 host logic -- shard plan, record exchange and merge algebra.

Each rank builds the record of its K-shard with the oracle (checker), exchanges it through the product's
host exchange (motion_planning_b200.distributed.exchange_host over gloo) and merges with the product's
merge statement; the merged update must equal the single-process update_action of the oracle
(control/src/mppi:186-196) -- i.e. sharding K changes nothing but the summation order.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from motion_planning_b200.distributed import exchange_host, merge_records, shard_plan
from oracle import mppi_oracle as orc


def shard_record(p, V, eps, lo, hi):
    """record (T,6) of rollouts [lo,hi): m, S, N0, N1, E0, E1 (layout of include/mppi_b200.h mppi_step_local)."""
    v = V[:, lo:hi]
    m = v.min(axis=1)
    e = np.exp(-(v - m[:, None]) / p.lam)
    rec = np.empty((p.T, 6))
    rec[:, 0] = m
    rec[:, 1] = e.sum(axis=1)
    rec[:, 2] = (e * eps[:, 0, lo:hi]).sum(axis=1)
    rec[:, 3] = (e * eps[:, 1, lo:hi]).sum(axis=1)
    rec[:, 4] = eps[:, 0, lo:hi].sum(axis=1)
    rec[:, 5] = eps[:, 1, lo:hi].sum(axis=1)
    return rec


def reference_dU(p, V, eps):
    """eps[t] @ (omega / sum omega) with omega = exp(-(V-min)/lam) + 1e-8  (control/src/mppi:189-196)."""
    dU = np.empty((2, p.T))
    for t in range(p.T):
        w = np.exp(-(V[t] - V[t].min()) / p.lam) + p.eps_floor
        dU[:, t] = eps[t] @ (w / w.sum())
    return dU


def _worker(rank, world, port, K, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = orc.Params(K=K, T=T)
        rng = np.random.RandomState(42)                    # same on every rank: the GLOBAL problem
        eps = rng.normal(0, 0.9, size=(T, 2, K))
        U = rng.normal(size=(2, T))
        x0, goal = np.array([0.1, 0.0, 0.2]), np.array([0.0, -1.0, 0.0])
        V = orc.get_cost2go(p, x0, U, goal, eps)
        kl, ko = shard_plan(K, world, rank)
        rec = shard_record(p, V, eps, ko, ko + kl)
        allrec = exchange_host(torch, dist, rec.reshape(-1), world).reshape(world, T, 6)
        assert np.array_equal(allrec[rank], rec)
        dU = merge_records(allrec, p.lam, p.eps_floor, K)
        want = reference_dU(p, V, eps)
        err = float(np.max(np.abs(dU - want)) / np.max(np.abs(want)))
        q.put((rank, err, kl, ko))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K,T", [(1000, 16), (257, 8)])
@pytest.mark.timeout(300)
def test_two_rank_shard_exchange_merge(K, T):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = tmp.get_context("fork")       # fork: the children inherit the already-imported torch (fast)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, T, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for _, err, kl, ko in res:
        assert err < 1e-12
    assert sum(r[2] for r in res) == K


def test_shard_plan_covers_everything_once():
    for K in (1, 7, 64, 65536, 2097152, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_plan(K, world, r) for r in range(world)]
            assert sum(k for k, _ in spans) == K
            off = 0
            for k, o in spans:
                assert o == off and k >= 0
                off += k
            assert max(k for k, _ in spans) - min(k for k, _ in spans) <= 1


def test_merge_is_invariant_to_the_number_of_shards():
    K, T = 4096, 12
    p = orc.Params(K=K, T=T)
    rng = np.random.RandomState(0)
    eps = rng.normal(0, 0.9, size=(T, 2, K))
    V = 3e4 + rng.gamma(2.0, 0.5, size=(T, K))          # spread >> lam: an arg-min-like softmin
    want = reference_dU(p, V, eps)
    for world in (1, 2, 5, 8):
        recs = []
        for r in range(world):
            kl, ko = shard_plan(K, world, r)
            recs.append(shard_record(p, V, eps, ko, ko + kl))
        dU = merge_records(np.stack(recs), p.lam, p.eps_floor, K)
        np.testing.assert_allclose(dU, want, rtol=1e-12, atol=1e-14)


# ---- the split-phase step's retry round trip (MPPI_ERR_RETRY) through ShardedMPPI.get_path, without a GPU ----------------
class _StubLib(object):
    """Stands in for libmppi_b200.so behind ShardedMPPI.get_path (exchange='host'): every rank's record carries its rank and
    the attempt number; mppi_step_finish answers MPPI_ERR_RETRY for the first attempt of step 1 -- on EVERY rank, because the
    decision is a function of the gathered records only (as in finalize_body, csrc/reduce_kernels.cuh)."""

    def __init__(self, rank, world, T):
        self.rank, self.world, self.T = rank, world, T
        self.step, self.attempt, self.gathered, self.log = 0, 0, None, []

    def mppi_set_goal(self, h, g):
        return 0

    def mppi_step_local(self, h, x):
        self.log.append(("local", self.step, self.attempt))
        return 0

    def mppi_read_record(self, h, rec):
        a = np.ctypeslib.as_array(rec, shape=(self.T * 6,))
        a[:] = 100.0 * self.rank + 10.0 * self.step + self.attempt
        if self.step == 1 and self.attempt == 0 and self.rank == 1:
            a[1] = -1.0                                     # rank 1's screen overflowed: S < 0 marks it for everybody
        return 0

    def mppi_write_gather(self, h, allrec):
        self.gathered = np.ctypeslib.as_array(allrec, shape=(self.world, self.T * 6)).copy()
        return 0

    def mppi_step_finish(self, h, u, x):
        overflow = bool((self.gathered[:, 1] < 0).any())
        self.log.append(("finish", self.step, self.attempt, overflow))
        if overflow:
            self.attempt += 1
            return 7                                        # MPPI_ERR_RETRY
        np.ctypeslib.as_array(x, shape=(3,))[:] = self.gathered[:, 0].sum()
        np.ctypeslib.as_array(u, shape=(2,))[:] = self.step
        self.step, self.attempt = self.step + 1, 0
        return 0

    def mppi_last_error(self):
        return b"stub"


def _retry_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from motion_planning_b200 import distributed as D
        T = 8
        sh = object.__new__(D.ShardedMPPI)
        sh._torch, sh._dist, sh.group, sh.world, sh.rank, sh.exchange, sh._n_rec = torch, dist, None, world, rank, "host", T * 6
        lib = _StubLib(rank, world, T)

        class _M(object):
            _lib, _h, dt, fin_time = lib, None, 1.0 / T, [0]
            _path_log, _uvec_log = [], []

            def _sync_sampling(self, sig, lam):
                pass
        sh.mppi = _M()
        outs = [sh.get_path(np.zeros(3), np.zeros(3)) for _ in range(3)]
        q.put((rank, [float(o[0]) for o in outs], lib.log))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_get_path_repeats_the_round_trip_on_retry():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = tmp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_retry_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (r0, out0, log0), (r1, out1, log1) = res
    assert out0 == out1                                      # every rank ends every step with the same result
    assert out0 == [100.0, 100.0 + 20.0 + 2.0, 100.0 + 40.0]  # step 1 was finished by its SECOND attempt (attempt digit 1 per rank)
    for log in (log0, log1):                                  # both ranks: local/finish, local/finish(retry)/local/finish, local/finish
        assert [e[0] for e in log] == ["local", "finish", "local", "finish", "local", "finish", "local", "finish"]
        assert [e[3] for e in log if e[0] == "finish"] == [False, True, False, False]
