"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the shard layout, the collectives
wrapper and the request/response routing of a sharded resample, with NumPy standing in for the device kernels
(the same ``route_resample`` code path the CUDA engine runs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qinfer_b200.sharded import (ShardComm, ShardLayout, cdf_bounds, exchange_plan, parity_resample, route_resample,
                                 split_counts, split_resample)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class NumpyOps(object):
    """CPU stand-in for the device kernels (qb_shard_classify / qb_shard_bucket / qb_draw / qb_gather_rows)."""

    def __init__(self, x_local, w_local):
        self.x = x_local
        self.cdf = np.cumsum(w_local)

    def classify(self, u, bounds):
        un = u.numpy()
        owner = np.searchsorted(np.asarray(bounds[1:-1]), un, side='right').astype(np.int32)
        counts = np.bincount(owner, minlength=len(bounds) - 1)
        return torch.from_numpy(owner), [int(c) for c in counts]

    def bucket(self, u, owner, bounds, starts):
        un, on = u.numpy(), owner.numpy()
        order = np.argsort(on, kind='stable')
        perm = np.empty_like(order)
        perm[order] = np.arange(order.size)
        req = (un - np.asarray(bounds)[on])[order]
        return torch.from_numpy(req), torch.from_numpy(perm.astype(np.int64))

    def local_draw(self, req_in):
        js = np.minimum(self.cdf.searchsorted(req_in.numpy(), side='right'), self.cdf.size - 1)
        return torch.from_numpy(js.astype(np.int64))

    def gather_rows(self, js):
        return torch.from_numpy(np.ascontiguousarray(self.x[js.numpy()]).reshape(-1))


def _worker(rank, world, port, n_global, d, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ShardComm()
        layout = ShardLayout(n_global, world)
        rs = np.random.RandomState(123)                    # every rank builds the same global problem
        x = rs.random_sample((n_global, d))
        w = rs.random_sample(n_global) ** 3
        w /= w.sum()
        u = rs.random_sample(n_global)
        lo, hi = layout.offsets[rank], layout.offsets[rank + 1]
        ops = NumpyOps(x[lo:hi], w[lo:hi])
        rows, perm, bounds = route_resample(comm, ops, torch.from_numpy(u[lo:hi].copy()), float(w[lo:hi].sum()), d)
        got = rows.numpy().reshape(-1, d)[perm.numpy()]
        # collectives wrapper
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        comm.all_reduce_sum(t)
        scal = comm.all_gather_scalars(rank * 2.5, torch.device("cpu"))
        cnt = comm.exchange_counts([rank * 10 + r for r in range(world)], torch.device("cpu"))
        objs = comm.all_gather_object(("rank", rank))
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), got=got, bounds=np.asarray(bounds), allred=t.numpy(),
                 scal=np.asarray(scal), cnt=np.asarray(cnt), nobj=len(objs))
        comm.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global,d", [(2, 1001, 1), (2, 4096, 3), (3, 500, 16)])
def test_route_resample_matches_single_process(tmp_path, world, n_global, d):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_global, d, str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(123)
    x = rs.random_sample((n_global, d))
    w = rs.random_sample(n_global) ** 3
    w /= w.sum()
    u = rs.random_sample(n_global)
    layout = ShardLayout(n_global, world)
    # single-process ground truth with the same shard-offset CDF the engine uses
    totals = [w[layout.offsets[r]:layout.offsets[r + 1]].sum() for r in range(world)]
    bounds = cdf_bounds(totals)
    global_cdf = np.concatenate([bounds[r] + np.cumsum(w[layout.offsets[r]:layout.offsets[r + 1]])
                                 for r in range(world)])
    plain = np.cumsum(w)
    assert np.max(np.abs(global_cdf - plain)) < 1e-13
    for r in range(world):
        f = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        lo, hi = layout.offsets[r], layout.offsets[r + 1]
        # what the routing must return: the owner-local bisection of (u - bounds[owner])
        owner = np.searchsorted(np.asarray(bounds[1:-1]), u[lo:hi], side='right')
        want = np.empty((hi - lo, d))
        for i, (ui, o) in enumerate(zip(u[lo:hi], owner)):
            olo, ohi = layout.offsets[o], layout.offsets[o + 1]
            j = min(np.cumsum(w[olo:ohi]).searchsorted(ui - bounds[o], side='right'), ohi - olo - 1)
            want[i] = x[olo + j]
        assert np.array_equal(f["got"], want)
        # and that equals the plain global draw except where u sits within rounding of a CDF step
        js_plain = np.minimum(plain.searchsorted(u[lo:hi], side='right'), n_global - 1)
        assert np.mean(np.all(f["got"] == x[js_plain], axis=1)) > 0.995
        assert np.allclose(f["bounds"], bounds)
        assert f["allred"][0] == sum(range(1, world + 1))
        assert list(f["scal"]) == [2.5 * q for q in range(world)]
        assert list(f["cnt"]) == [q * 10 + r for q in range(world)]
        assert int(f["nobj"]) == world


def test_shard_layout():
    L = ShardLayout(10, 3)
    assert L.counts == [4, 3, 3] and L.offsets == [0, 4, 7, 10]
    assert [L.owner_of_index(i) for i in (0, 3, 4, 6, 7, 9)] == [0, 0, 1, 1, 2, 2]
    assert ShardLayout(10 ** 8, 8).counts == [12500000] * 8
    with pytest.raises(ValueError):
        ShardLayout(2, 3)
    assert cdf_bounds([0.25, 0.5, 0.25]) == [0.0, 0.25, 0.75, 1.0]


# ---------------------------------------------------------------------------
# "split" resample: shared multinomial split + local draws + one all-to-all of the surplus rows
# ---------------------------------------------------------------------------
def test_exchange_plan_balances_every_rank():
    rs = np.random.RandomState(3)
    for G in (1, 2, 3, 8):
        cap = ShardLayout(1000 * G + 3, G).counts
        for _ in range(20):
            m = [int(c) for c in rs.multinomial(sum(cap), rs.dirichlet(np.ones(G) * 0.3))]
            T = np.array(exchange_plan(m, cap))
            assert T.shape == (G, G) and (T >= 0).all() and np.trace(T) == 0
            sent, recv = T.sum(axis=1), T.sum(axis=0)
            for r in range(G):
                assert sent[r] == max(m[r] - cap[r], 0)
                assert recv[r] == max(cap[r] - m[r], 0)
                assert min(m[r], cap[r]) + recv[r] == cap[r]
    assert exchange_plan([5, 5], [5, 5]) == [[0, 0], [0, 0]]
    assert exchange_plan([10, 0, 2], [4, 4, 4]) == [[0, 4, 2], [0, 0, 0], [0, 0, 0]]


def test_split_counts_is_a_shared_multinomial():
    masses = [0.1, 0.0, 0.6, 0.3]
    a = split_counts(np.random.Generator(np.random.Philox(key=7)), 10 ** 6, masses)
    b = split_counts(np.random.Generator(np.random.Philox(key=7)), 10 ** 6, masses)
    assert a == b and sum(a) == 10 ** 6 and a[1] == 0
    assert all(abs(c - 1e6 * p) < 6 * np.sqrt(1e6 * p * (1 - p)) + 1 for c, p in zip(a, masses))
    # degenerate masses fall back to an even split instead of raising
    c = split_counts(np.random.Generator(np.random.Philox(key=1)), 1000, [float('nan'), 0.0])
    assert sum(c) == 1000


class SplitOps(object):
    """CPU stand-in for the fused draw+move kernel: offspring rows are tagged (origin rank, serial number)."""

    def __init__(self, rank, cap, width):
        self.rank, self.width = rank, width
        self._slab = torch.full((cap, width), -1.0, dtype=torch.float64)
        self.made = 0

    def slab(self):
        return self._slab

    def alloc(self, rows):
        return torch.empty((rows, self.width), dtype=torch.float64)

    def draw_into(self, dst):
        k = dst.shape[0]
        assert k <= self._slab.shape[0]                 # chunks never exceed the slab (fixed scratch)
        dst[:, 0] = self.rank
        dst[:, 1:] = torch.arange(self.made, self.made + k, dtype=torch.float64)[:, None]
        self.made += k


def _split_worker(rank, world, port, n_global, masses, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ShardComm()
        layout = ShardLayout(n_global, world)
        rng = np.random.Generator(np.random.Philox(key=99))      # same key on every rank
        m = split_counts(rng, n_global, masses)
        ops = SplitOps(rank, layout.counts[rank], 3)
        sent, recv = split_resample(comm, ops, m, layout.counts, 3)
        rows = comm.all_gather_rows(torch.tensor([float(rank), 1.0, 2.0], dtype=torch.float64))
        np.savez(os.path.join(out_dir, "s%d.npz" % rank), slab=ops.slab().numpy(), m=np.asarray(m), made=ops.made,
                 sent=sent, recv=recv, rows=rows)
        comm.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global,masses", [(2, 1001, [0.5, 0.5]), (3, 3000, [0.7, 0.1, 0.2]),
                                                   (2, 400, [1.0, 0.0])])
def test_split_resample_moves_only_the_surplus(tmp_path, world, n_global, masses):
    port = _free_port()
    mp.spawn(_split_worker, args=(world, port, n_global, masses, str(tmp_path)), nprocs=world, join=True)
    layout = ShardLayout(n_global, world)
    files = [np.load(os.path.join(str(tmp_path), "s%d.npz" % r)) for r in range(world)]
    m = list(files[0]["m"])
    assert sum(m) == n_global
    seen = set()
    for r, f in enumerate(files):
        assert list(f["m"]) == m                              # every rank computed the same split
        assert int(f["made"]) == m[r]                         # and drew exactly its share from its own slab
        assert int(f["sent"]) == max(m[r] - layout.counts[r], 0)
        assert int(f["recv"]) == max(layout.counts[r] - m[r], 0)
        slab = f["slab"]
        assert slab.shape == (layout.counts[r], 3) and (slab[:, 0] >= 0).all()      # full, nothing left unwritten
        assert np.array_equal(slab[:, 1], slab[:, 2])
        keep = min(m[r], layout.counts[r])
        assert (slab[:keep, 0] == r).all()                    # own offspring stay in place
        seen.update((int(o), int(k)) for o, k in slab[:, :2])
        assert np.array_equal(f["rows"], np.array([[q, 1.0, 2.0] for q in range(world)]))
    # the union of the slabs is exactly the set of offspring drawn: nothing lost, nothing duplicated
    assert seen == {(r, k) for r in range(world) for k in range(m[r])}


# ---------------------------------------------------------------------------------------------------------------------
# parity mode: chained sequential scan + shared legacy stream + the reference's `mus[:k]` prefix in global
# invalid order.  NumPy stands in for the kernels; the ORACLE's LiuWestResampler on the whole cloud is the truth.
# ---------------------------------------------------------------------------------------------------------------------
class NumpyParityOps(object):
    def __init__(self, model, x_local, w_local, mean, S, a):
        self.model, self.x, self.w, self.mean, self.S, self.a = model, x_local, w_local, mean, S, a
        self.new = None
        self.invalid = None

    def zeros(self, n):
        return torch.zeros((n,), dtype=torch.float64)

    def slab(self):
        return torch.from_numpy(self.x)

    def scan(self, carry):
        run = float(carry[0])
        out = np.empty_like(self.w)
        for i, wi in enumerate(self.w):               # the sequential fp64 chain, continued from the previous slab
            run = run + wi
            out[i] = run
        return torch.from_numpy(out)

    def uniforms(self, n):
        return torch.from_numpy(np.random.random((n,)))

    def normals(self, d, k):
        return torch.from_numpy(np.random.randn(d, k).reshape(-1))

    def draw(self, cdf, u):
        c = cdf.numpy()
        return torch.from_numpy(np.minimum(c.searchsorted(u.numpy(), side='right'), c.size - 1).astype(np.int64))

    def _perturb(self, x_all, js, eps):
        d = self.x.shape[1]
        mus = self.a * x_all.numpy()[js.numpy(), :] + (1 - self.a) * self.mean
        return mus + np.dot(self.S, eps.numpy().reshape(d, -1)).T

    def move(self, x_all, js, eps):
        self.new = self._perturb(x_all, js, eps)
        self.invalid = np.logical_not(self.model.are_models_valid(self.new))
        return int(self.invalid.sum())

    def retry(self, x_all, js_prefix, eps, k):
        idxs = np.nonzero(self.invalid)[0]
        assert idxs.size == k
        self.new[idxs] = self._perturb(x_all, js_prefix, eps)
        self.invalid[idxs] = np.logical_not(self.model.are_models_valid(self.new[idxs]))
        return int(self.invalid.sum())


def _parity_problem(n_global, d, seed):
    import smc_oracle as oracle
    rs = np.random.RandomState(seed)
    if d == 1:
        model = oracle.SimplePrecessionModel(min_freq=0.35)            # a third of the perturbed cloud is invalid
        x = 0.3 + 0.4 * rs.random_sample((n_global, 1))
    else:
        model = oracle.RandomizedBenchmarkingModel()                   # p, A, B in [0, 1], A + B <= 1
        x = np.column_stack([0.9 + 0.1 * rs.random_sample(n_global), 0.6 * rs.random_sample(n_global),
                             0.4 * rs.random_sample(n_global)])
    w = rs.random_sample(n_global) ** 4
    w[rs.randint(0, n_global, n_global // 7)] = 0.0
    w /= w.sum()
    return oracle, model, x, w


def _parity_worker(rank, world, port, n_global, d, out_dir):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle, model, x, w = _parity_problem(n_global, d, 321)
        comm = ShardComm()
        layout = ShardLayout(n_global, world)
        lo, hi = layout.offsets[rank], layout.offsets[rank + 1]
        a, h = 0.9, np.sqrt(1 - 0.9 ** 2)
        whole = oracle.ParticleDistribution(particle_locations=x, particle_weights=w)
        mean = whole.est_mean()
        S = np.real(h * oracle.sqrtm_psd(whole.est_covariance_mtx())[0])
        np.random.seed(77)                                              # the same legacy stream on every rank
        ops = NumpyParityOps(model, x[lo:hi].copy(), w[lo:hi].copy(), mean, S, a)
        n_iters, bad, js = parity_resample(comm, ops, layout, d, True, 1000)
        np.savez(os.path.join(out_dir, "p%d.npz" % rank), new=ops.new, js=js.numpy(), n_iters=n_iters, bad=bad,
                 next_u=np.random.random())
        comm.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global,d", [(2, 3001, 1), (3, 2000, 3)])
def test_parity_resample_equals_the_single_process_reference_algorithm(tmp_path, world, n_global, d):
    port = _free_port()
    mp.spawn(_parity_worker, args=(world, port, n_global, d, str(tmp_path)), nprocs=world, join=True)
    oracle, model, x, w = _parity_problem(n_global, d, 321)
    np.random.seed(77)
    dist_ = oracle.ParticleDistribution(particle_locations=x, particle_weights=w)
    res = oracle.LiuWestResampler(a=0.9)
    want = res(model, dist_).particle_locations
    next_u = np.random.random()
    js_want = np.cumsum(w).searchsorted(np.random.RandomState(77).random_sample(n_global), side='right')
    parts = [np.load(os.path.join(str(tmp_path), "p%d.npz" % r)) for r in range(world)]
    got = np.concatenate([f["new"] for f in parts], axis=0)
    assert np.array_equal(np.concatenate([f["js"] for f in parts]), js_want)        # global indices, bit for bit
    assert all(int(f["n_iters"]) > 2 and int(f["bad"]) == 0 for f in parts)         # the retry loop really ran
    assert all(float(f["next_u"]) == next_u for f in parts)                         # streams consumed in lockstep
    if d == 1:
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)


def _comm_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ShardComm()
        cpu = torch.device("cpu")
        # ragged all-gather of slabs of rows (width 3, unequal counts), in rank order
        counts = [5, 2, 4][:world]
        mine = torch.arange(counts[rank] * 3, dtype=torch.float64) + 100.0 * rank
        ragged = comm.all_gather_ragged(mine, counts, 3)
        seed = comm.broadcast_int(12345 + rank, cpu)                 # rank 0's value everywhere
        ints = comm.all_gather_ints(7 * rank + 1, cpu)
        # the chain: a running value travels rank to rank
        carry = torch.zeros((1,), dtype=torch.float64)
        if rank > 0:
            comm.recv_prev(carry)
        carry = carry + float(rank + 1)
        if rank < world - 1:
            comm.send_next(carry)
        np.savez(os.path.join(out_dir, "c%d.npz" % rank), ragged=ragged.numpy(), seed=seed, ints=np.asarray(ints),
                 carry=carry.numpy())
        comm.barrier()
    finally:
        dist.destroy_process_group()


def test_comm_helpers_of_the_parity_mode(tmp_path):
    world = 3
    mp.spawn(_comm_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    counts = [5, 2, 4]
    want = np.concatenate([np.arange(counts[r] * 3, dtype=np.float64) + 100.0 * r for r in range(world)])
    for r in range(world):
        f = np.load(os.path.join(str(tmp_path), "c%d.npz" % r))
        assert np.array_equal(f["ragged"], want)
        assert int(f["seed"]) == 12345
        assert list(f["ints"]) == [1, 8, 15]
        assert float(f["carry"][0]) == sum(range(1, r + 2))          # 1, 1 + 2, 1 + 2 + 3
