"""Sharded particle cloud: one contiguous slab of particles per GPU.

The reference's only multi-process strategy splits the particle axis of
``likelihood`` across ipyparallel engines (parallel.py:183-224).  Here the whole
updater is sharded the same way, one process per GPU:

  update    each rank runs the fused kernel on its slab; the three sums it needs globally (sum w', sum w'^2,
            #bad) are all-reduced INSIDE that launch over the peers' NVLink-mapped mailboxes (qb_update.cu),
            so every rank publishes bit-identical global stats with no NCCL call and no extra launch.
  resample  d <= 4 ("split", the default): ONE all-gather of every rank's 1 + d + d^2 moment sums gives the global
            mean/covariance AND the shard masses; every rank then draws the same multinomial split m ~ Mult(N, masses)
            from a shared counter-based host generator, draws m[r] offspring from ITS OWN slab with the fused
            draw+move kernel (a multinomial draw of N from the global CDF is exactly: split the count over the
            shards, then draw within each shard), and only the surplus |m[r] - N/G| rows travel, in ONE all-to-all
            whose counts every rank already knows (no count exchange, no index routing).
            d > 4 ("route"): all-reduce of the moment sums, all-gather of the G shard totals (the global CDF is the
            shard CDFs offset by their exclusive prefix), then ONE request/response all-to-all: every draw is
            classified to the shard that owns its CDF range, the owner bisects its local CDF and sends the
            row back; shrink + perturb + validity run locally (NCCL through torch.distributed).

``ShardComm`` is the only place that touches torch.distributed, and the routing of a resample is written against
a small ``ops`` interface so that it runs unchanged on CPU tensors under gloo (tests/test_sharded_cpu.py).

  parity    ``LiuWestResampler(rng='numpy' | 'mt19937')`` (SURVEY §8e "Parity mode"): the global resample indices
            equal the single-process reference's.  The exact scan is CHAINED across the slabs — rank r continues the
            sequential fp64 sum from the last CDF entry of rank r - 1 (one double, point to point) — the slab CDFs
            and slabs are all-gathered, every rank generates the same legacy stream (a sequential generator cannot be
            sharded) and keeps the variates of its own global slots, and the postselection retry follows the
            reference's ``mus[:k]`` prefix (resamplers.py:372) in GLOBAL invalid order (``parity_resample`` below).

Deviations of the throughput mode from the single-GPU path, both documented in DESIGN.md: the CDF is the
re-associated (fast) scan, so resample indices are not bit-identical to the single-process reference; and the
postselection retry re-centres a particle on its OWN shrunk mean instead of replicating the reference's prefix-slice
quirk (resamplers.py:372), which would need a second exchange.
"""
import ctypes
import warnings

import numpy as np
import scipy.linalg
import torch

from . import _lib
from ._lib import nvtx_range
from ._exceptions import ResamplerError, ResamplerWarning


# ---------------------------------------------------------------------------
# layout + communication (host logic, runs on CPU under gloo as well)
# ---------------------------------------------------------------------------
class ShardLayout(object):
    """Contiguous slabs: global index = offsets[rank] + local index (parallel.py:218 uses the same split)."""

    def __init__(self, n_global, world):
        if n_global < world:
            raise ValueError("need at least one particle per rank")
        base, extra = divmod(int(n_global), int(world))
        self.n_global = int(n_global)
        self.world = int(world)
        self.counts = [base + (1 if r < extra else 0) for r in range(world)]
        self.offsets = [0]
        for c in self.counts:
            self.offsets.append(self.offsets[-1] + c)

    def count(self, rank):
        return self.counts[rank]

    def set_counts(self, counts):
        """Floating slabs: after a resample every rank keeps the offspring it drew (counts sum to n_global)."""
        assert sum(counts) == self.n_global and len(counts) == self.world
        self.counts = [int(c) for c in counts]
        self.offsets = [0]
        for c in self.counts:
            self.offsets.append(self.offsets[-1] + c)

    def owner_of_index(self, global_index):
        return int(np.searchsorted(self.offsets, global_index, side='right') - 1)


def cdf_bounds(shard_totals):
    """bounds[r] = global CDF value at the start of shard r; bounds[G] = total (fixed rank order => identical
    on every rank)."""
    b = [0.0]
    for t in shard_totals:
        b.append(b[-1] + float(t))
    return b


class ShardComm(object):
    """torch.distributed plumbing.  Works on CUDA tensors with NCCL and on CPU tensors with gloo."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_scalars(self, value, device):
        """[value of rank 0, ..., value of rank G-1] as Python floats (one collective, one host read)."""
        mine = torch.tensor([float(value)], dtype=torch.float64, device=device)
        out = torch.empty((self.world,), dtype=torch.float64, device=device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return [float(v) for v in out.tolist()]

    def exchange_counts(self, send_counts, device):
        """recv[r] = how many items rank r sends me.  One all-gather of the G x G count matrix (a single
        collective and a single host read) instead of an all-to-all."""
        send = torch.tensor([int(c) for c in send_counts], dtype=torch.int64, device=device)
        mat = torch.empty((self.world * self.world,), dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(mat, send, group=self.group)
        m = mat.tolist()
        return [int(m[r * self.world + self.rank]) for r in range(self.world)]

    def all_to_all_v(self, send, send_counts, recv_counts, width=1):
        """Variable all-to-all of rows of ``width`` elements; ``send`` is bucketed by destination rank."""
        recv = torch.empty((int(sum(recv_counts)) * width,), dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(recv, send.reshape(-1),
                                    output_split_sizes=[int(c) * width for c in recv_counts],
                                    input_split_sizes=[int(c) * width for c in send_counts], group=self.group)
        return recv

    def all_gather_rows_begin(self, row):
        """Queue the all-gather of every rank's ``row`` and its read-back; ``all_gather_rows_end`` waits for exactly
        that (an event), so kernels queued in between keep the device busy meanwhile."""
        out = torch.empty((self.world * row.numel(),), dtype=row.dtype, device=row.device)
        self.dist.all_gather_into_tensor(out, row.contiguous(), group=self.group)
        if out.is_cuda:
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return host, ev, row.numel(), out
        return out, None, row.numel(), out

    def all_gather_rows_end(self, handle):
        host, ev, width, _keep = handle
        if ev is not None:
            ev.synchronize()
        return host.numpy().reshape(self.world, width).copy()

    def all_gather_rows(self, row):
        """(world, len(row)) host array of every rank's ``row`` (one collective, one host read)."""
        return self.all_gather_rows_end(self.all_gather_rows_begin(row))

    def all_to_all_into(self, recv, send, send_counts, recv_counts, width=1):
        """Variable all-to-all of rows of ``width`` elements straight into ``recv`` (a flat, contiguous view)."""
        self.dist.all_to_all_single(recv, send.reshape(-1),
                                    output_split_sizes=[int(c) * width for c in recv_counts],
                                    input_split_sizes=[int(c) * width for c in send_counts], group=self.group)
        return recv

    def _global_rank(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def recv_prev(self, t):
        """Point to point: receive ``t`` from rank - 1 (the carried-in running sum of a chained scan)."""
        self.dist.recv(t, src=self._global_rank(self.rank - 1), group=self.group)
        return t

    def send_next(self, t):
        self.dist.send(t.contiguous(), dst=self._global_rank(self.rank + 1), group=self.group)

    def all_gather_ragged(self, t, counts, width=1):
        """Concatenation over the ranks of slabs of counts[r] rows of ``width`` elements (flat tensors)."""
        counts = [int(c) for c in counts]
        cap = max(counts) * width
        mine = t.reshape(-1)
        if mine.numel() != cap:
            pad = torch.zeros((cap,), dtype=t.dtype, device=t.device)
            pad[:mine.numel()] = mine
            mine = pad
        out = torch.empty((self.world * cap,), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, mine.contiguous(), group=self.group)
        if all(c * width == cap for c in counts):
            return out
        return torch.cat([out[q * cap:q * cap + counts[q] * width] for q in range(self.world)])

    def all_gather_ints(self, value, device):
        mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
        out = torch.empty((self.world,), dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return [int(v) for v in out.tolist()]

    def broadcast_int(self, value, device):
        """Rank 0's ``value`` (a non-negative int below 2^63) on every rank: one small collective."""
        t = torch.tensor([int(value)], dtype=torch.int64, device=device)
        self.dist.broadcast(t, src=self._global_rank(0), group=self.group)
        return int(t.item())

    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)


def route_resample(comm, ops, u, shard_total, width):
    """The request/response exchange of one sharded multinomial draw.

    ``u``: this rank's uniforms (one per particle it will own after the resample).  ``shard_total``: the sum of
    this rank's normalised weights.  ``ops`` supplies the local arithmetic (CUDA kernels in the product, NumPy in
    the CPU tests): classify / bucket / local_draw / gather_rows.  Returns ``(rows, perm)``: ``rows[perm[i]]`` is
    the old particle drawn for slot ``i`` (``rows`` is (n, width) flattened).
    """
    device = u.device
    bounds = cdf_bounds(comm.all_gather_scalars(shard_total, device))
    owner, counts = ops.classify(u, bounds)                      # counts[r]: how many of my draws shard r owns
    starts = [0]
    for c in counts[:-1]:
        starts.append(starts[-1] + int(c))
    req, perm = ops.bucket(u, owner, bounds, starts)             # owner-local CDF coordinates, bucketed by owner
    recv_counts = comm.exchange_counts(counts, device)
    req_in = comm.all_to_all_v(req, counts, recv_counts, 1)
    js = ops.local_draw(req_in)                                  # bisect my local CDF for everybody's requests
    rows_out = ops.gather_rows(js)
    rows = comm.all_to_all_v(rows_out, recv_counts, counts, width)
    return rows, perm, bounds


def split_counts(rng, n_global, shard_masses):
    """How many of the ``n_global`` offspring each shard produces: one multinomial draw over the shard masses.
    ``rng`` is a counter-based generator seeded identically on every rank, so all ranks compute the same split
    without communicating.  (Drawing N indices from the global CDF and counting them per shard has exactly this
    distribution; the draws inside a shard are then i.i.d. from that shard's normalised weights.)"""
    p = np.asarray(shard_masses, dtype=np.float64).copy()
    p[~np.isfinite(p) | (p < 0)] = 0.0
    if p.sum() <= 0:
        p[:] = 1.0
    p /= p.sum()
    p[-1] = max(0.0, 1.0 - p[:-1].sum())         # Generator.multinomial wants sum(p[:-1]) <= 1
    return [int(c) for c in rng.multinomial(int(n_global), p)]


def exchange_plan(m, cap):
    """T[r][q] = rows rank r sends to rank q so that every rank ends with cap[r] rows: surplus ranks hand their
    extra offspring to deficit ranks, both visited in rank order (deterministic, computed by every rank)."""
    G = len(m)
    assert sum(m) == sum(cap)
    T = [[0] * G for _ in range(G)]
    need = [cap[r] - m[r] for r in range(G)]       # > 0: deficit
    q = 0
    for r in range(G):
        s = m[r] - cap[r]
        while s > 0:
            while need[q] <= 0:
                q += 1
            t = min(s, need[q])
            T[r][q] += t
            need[q] -= t
            s -= t
    return T


def split_resample(comm, ops, m, cap, width):
    """The exchange of one "split" resample.  ``ops.slab()`` is this rank's (cap, width) destination,
    ``ops.draw_into(dst2d)`` fills ``dst2d`` with offspring of the local slab, ``ops.alloc(rows)`` returns a
    (rows, width) send buffer.  Returns how many rows this rank sent and received."""
    r = comm.rank
    keep = min(m[r], cap[r])
    slab = ops.slab()
    if keep:
        ops.draw_into(slab[:keep])
    extra = m[r] - keep
    send = ops.alloc(extra)
    done = 0
    while done < extra:                                # chunks no larger than the slab (fixed scratch)
        nc = min(cap[r], extra - done)
        ops.draw_into(send[done:done + nc])
        done += nc
    T = exchange_plan(m, cap)
    if any(any(row) for row in T):
        recv_counts = [T[q][r] for q in range(comm.world)]
        assert sum(recv_counts) == cap[r] - keep
        comm.all_to_all_into(slab[keep:].reshape(-1), send, T[r], recv_counts, width)
    return extra, cap[r] - keep


# ---------------------------------------------------------------------------
# device ops + sharded updater
# ---------------------------------------------------------------------------
def _columns(eps, d, k_total, first, k):
    """Columns [first, first + k) of a flat (d, k_total) row-major block, as a flat (d, k) block."""
    if d == 1 or k == k_total:
        return eps[first:first + k] if d == 1 else eps
    return eps.reshape(d, k_total)[:, first:first + k].contiguous().reshape(-1)


def parity_resample(comm, ops, layout, d, postselect, maxiter):
    """Liu-West resample of a sharded cloud with the reference's GLOBAL indices (resamplers.py:308-372).

    ``ops`` supplies the kernels (CUDA engine, or NumPy in tests/test_sharded_cpu.py):
      scan(carry) -> this slab's piece of np.cumsum(w_global), continued from the one-element tensor ``carry``
      slab() -> (n_local, d) locations;  zeros(n) -> float64 tensor
      uniforms(n) / normals(d, k) -> the next n / d*k variates of the legacy stream (identical on every rank)
      draw(cdf, u) -> min(searchsorted(cdf, u, 'right'), len(cdf) - 1)
      move(x_all, js, eps) -> writes the new slab, returns this rank's invalid count
      retry(x_all, js_prefix, eps, k) -> re-perturbs the k invalid rows (ascending slot order) around
                                         a x_all[js_prefix[r]] + (1 - a) mean, returns the new invalid count
    Returns (n_iters, n_invalid_global, js of this slab)."""
    r, world = comm.rank, comm.world
    counts, first = layout.counts, layout.offsets[comm.rank]
    n_local, n_global = counts[r], layout.n_global
    # the sequential-CDF chain crosses the shards: one double travels rank to rank
    carry = ops.zeros(1)
    if r > 0:
        comm.recv_prev(carry)
    cdf = ops.scan(carry)
    if r < world - 1:
        comm.send_next(cdf[n_local - 1:n_local])
    cdf_all = comm.all_gather_ragged(cdf, counts, 1)
    x_all = comm.all_gather_ragged(ops.slab(), counts, d).reshape(n_global, d)
    u_all = ops.uniforms(n_global)                                   # resamplers.py:319
    js = ops.draw(cdf_all, u_all[first:first + n_local])             # the global js of this rank's slots
    eps = ops.normals(d, n_global)                                   # kernel(n_rvs, n) — one global block
    k_local = ops.move(x_all, js, _columns(eps, d, n_global, first, n_local))
    n_iters = 1
    while True:
        ks = comm.all_gather_ints(k_local if postselect else 0, cdf.device)
        k_global = sum(ks)
        if k_global == 0 or n_iters >= maxiter:
            return n_iters, k_global, js
        n_iters += 1
        # `mus = mus[:k]` (resamplers.py:372): the q-th still-invalid particle IN GLOBAL ORDER is re-centred on the
        # q-th original draw; its parent index is recomputed here from the shared uniform stream
        q0 = sum(ks[:r])
        eps = ops.normals(d, k_global)                               # every rank consumes the whole block
        if k_local:
            js_prefix = ops.draw(cdf_all, u_all[q0:q0 + k_local])
            k_local = ops.retry(x_all, js_prefix, _columns(eps, d, k_global, q0, k_local), k_local)


class _ParityOps(object):
    """``parity_resample``'s kernels on the CUDA engine."""

    def __init__(self, cloud, resampler, mean, S):
        self.cloud, self.res, self.mean, self.S = cloud, resampler, mean, S
        cloud._resample_scratch(cloud.n)
        self._bufs = {}

    def _buf(self, name, n, dtype=torch.float64):
        b = self._bufs.get(name)
        if b is None or b.numel() < n:
            self._bufs[name] = b = torch.empty((int(n),), dtype=dtype, device=self.cloud.device)
        return b[:n]

    def zeros(self, n):
        return torch.zeros((n,), dtype=torch.float64, device=self.cloud.device)

    def slab(self):
        return self.cloud.x

    def scan(self, carry):
        return self.cloud.cdf(_lib.QB_SCAN_EXACT, carry=carry)

    def uniforms(self, n):
        out = self._buf('u', n)
        if self.res._rng == 'numpy':
            out.copy_(torch.from_numpy(np.random.random((n,))))
        else:
            self.cloud.mt19937_uniform(out, n)
        return out

    def normals(self, d, k):
        out = self._buf('eps', d * k)
        if self.res._rng == 'numpy':
            eps = np.ascontiguousarray(self.res._kernel(d, k), dtype=np.float64)
            if eps.shape != (d, k):
                raise ValueError("resampling kernel returned shape %s, expected %s" % (eps.shape, (d, k)))
            out.copy_(torch.from_numpy(eps.reshape(-1)))
        else:
            self.cloud.mt19937_normal(out, d * k)
        return out

    def draw(self, cdf, u):
        cloud = self.cloud
        js = torch.empty((u.numel(),), dtype=torch.int64, device=cloud.device)
        nbytes = cloud.lib.qb_draw_workspace_bytes(cdf.numel())
        ws = self._buf('draw_ws', max(nbytes, 8), torch.uint8)
        _lib.check(cloud.lib.qb_draw(ctypes.c_void_p(cdf.data_ptr()), cdf.numel(), ctypes.c_void_p(u.data_ptr()),
                                     u.numel(), ctypes.c_void_p(js.data_ptr()),
                                     ctypes.c_void_p(cloud.counter[1:].data_ptr()),
                                     ctypes.c_void_p(ws.data_ptr()) if nbytes else None, nbytes,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        cloud.launches += 2
        return js

    def move(self, x_all, js, eps):
        res = self.res
        self.cloud.lw_move(self.mean, self.S, res._a, eps, self.cloud.n, res._postselect, x_src=x_all, js=js)
        return self.cloud.read_counter()[0] if res._postselect else 0

    def retry(self, x_all, js_prefix, eps, k):
        self.cloud.compact_invalid(self.cloud.n)
        self.cloud.lw_retry(self.mean, self.S, self.res._a, eps, k, x_src=x_all, js=js_prefix)
        return self.cloud.read_counter()[0]


class _DeviceOps(object):
    def __init__(self, cloud, world):
        self.cloud = cloud
        self.world = world
        dev = cloud.device
        self.owner = torch.empty((cloud.n,), dtype=torch.int32, device=dev)
        self.counts = torch.zeros((world,), dtype=torch.int64, device=dev)
        self.cursor = torch.zeros((world,), dtype=torch.int64, device=dev)
        self.req = torch.empty((cloud.n,), dtype=torch.float64, device=dev)
        self.perm = torch.empty((cloud.n,), dtype=torch.int64, device=dev)
        self.overflow = torch.zeros((1,), dtype=torch.int64, device=dev)

    def classify(self, u, bounds):
        from .engine import _ptr, _stream
        c = self.cloud
        _lib.check(c.lib.qb_shard_classify(_ptr(u), u.numel(), _lib.f64_array(bounds), self.world, _ptr(self.owner),
                                           _ptr(self.counts), _stream()))
        c.launches += 1
        return self.owner, [int(v) for v in self.counts.tolist()]

    def bucket(self, u, owner, bounds, starts):
        from .engine import _ptr, _stream
        c = self.cloud
        st = (ctypes.c_int64 * self.world)(*[int(s) for s in starts])
        _lib.check(c.lib.qb_shard_bucket(_ptr(u), _ptr(owner), u.numel(), _lib.f64_array(bounds), st, self.world,
                                         _ptr(self.cursor), _ptr(self.req), _ptr(self.perm), _stream()))
        c.launches += 1
        return self.req[:u.numel()], self.perm[:u.numel()]

    def local_draw(self, req_in):
        from .engine import _ptr, _stream
        c = self.cloud
        m = req_in.numel()
        js = torch.empty((max(m, 1),), dtype=torch.int64, device=c.device)
        if m:
            _lib.check(c.lib.qb_draw(_ptr(c._cdf), c.n, _ptr(req_in), m, _ptr(js), _ptr(self.overflow), _ptr(c.ws),
                                     c.ws_bytes, _stream()))
            c.launches += 2
        return js[:m]

    def gather_rows(self, js):
        from .engine import _ptr, _stream
        c = self.cloud
        m = js.numel()
        out = torch.empty((max(m, 1) * c.d,), dtype=torch.float64, device=c.device)
        if m:
            _lib.check(c.lib.qb_gather_rows(_ptr(c.x), c.d, _ptr(js), m, _ptr(out), _stream()))
            c.launches += 1
        return out[:m * c.d]


_MAILBOX_CACHE = {}      # (group, world, rank, device) -> PeerMailboxes released by a closed updater
_WARMED = set()          # groups whose NCCL channels the resample exchange has already created


def _group_key(comm):
    return (id(comm.group) if comm.group is not None else 0, comm.world, comm.rank, torch.cuda.current_device())


class PeerMailboxes(object):
    """One mailbox per rank, mapped into every peer process through CUDA IPC.

    Creating one costs milliseconds (cudaMalloc, handle exchange, cudaIpcOpenMemHandle on every peer), so a closed
    updater RELEASES its mailboxes to a per-process cache and the next sharded updater of the same group reuses them.
    Rows are validated by the launch tag they carry: the new cloud continues the tag sequence where the previous one
    stopped (``last_tag``; identical on every rank, the ranks issue the same launches), so a stale row never matches."""

    @classmethod
    def acquire(cls, comm):
        mb = _MAILBOX_CACHE.pop(_group_key(comm), None)
        if mb is None:
            mb = cls(comm)
        mb.comm = comm
        return mb

    def release(self, last_tag):
        """Park the mapping for the next updater (every rank's launches have drained: synchronize + barrier)."""
        torch.cuda.synchronize()
        self.comm.barrier()
        self.last_tag = max(int(last_tag), self.last_tag)
        _MAILBOX_CACHE[_group_key(self.comm)] = self

    def __init__(self, comm):
        lib = _lib.load()
        self.lib = lib
        self.comm = comm
        self.last_tag = 0
        mine = ctypes.c_void_p()
        _lib.check(lib.qb_mailbox_create(comm.world, ctypes.byref(mine)))
        self.mine = mine
        handle = ctypes.create_string_buffer(_lib.QB_IPC_HANDLE_BYTES)
        _lib.check(lib.qb_ipc_get_handle(mine, handle))
        handles = comm.all_gather_object(bytes(handle.raw))
        self.ptrs = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == comm.rank:
                self.ptrs.append(mine.value)
            else:
                p = ctypes.c_void_p()
                _lib.check(lib.qb_ipc_open_handle(ctypes.create_string_buffer(h, _lib.QB_IPC_HANDLE_BYTES),
                                                  ctypes.byref(p)))
                self.ptrs.append(p.value)
                self._opened.append(p)
        self.error_flag = torch.zeros((1,), dtype=torch.int32, device='cuda')
        comm.barrier()

    def install(self, ctl):
        ctl.n_ranks = self.comm.world
        ctl.rank = self.comm.rank
        for r, p in enumerate(self.ptrs):
            ctl.d_peer_mailbox[r] = p
        ctl.d_error_flag = self.error_flag.data_ptr()

    def close(self):
        try:
            torch.cuda.synchronize()
            self.comm.barrier()
            for p in self._opened:
                self.lib.qb_ipc_close_handle(p)
            self._opened = []
            self.comm.barrier()
            if self.mine is not None:
                self.lib.qb_mailbox_destroy(self.mine)
                self.mine = None
        except Exception:  # pragma: no cover - interpreter shutdown
            pass


def _make_sharded_updater_class():
    from .smc import SMCUpdater
    from .distributions import covariance_from_moments
    from .resamplers import sqrtm_psd, _cov_1x1

    class ShardedSMCUpdater(SMCUpdater):
        """``SMCUpdater`` whose cloud is sharded over the ranks of the default process group.

        ``n_particles`` is the GLOBAL particle count; ``prior.sample(n)`` is called with this rank's slab size
        (seed the ranks differently).  ``particle_locations`` / ``particle_weights`` are this rank's slab (weights
        normalised globally); ``est_mean`` / ``est_covariance_mtx`` / ``n_ess`` / records are global and identical
        on every rank, so the control flow (resample decisions, policies) stays in lockstep without extra
        communication.
        """

        def __init__(self, model, n_particles, prior, group=None, exchange='split', **kwargs):
            if exchange not in ('split', 'route'):
                raise ValueError("exchange must be 'split' or 'route'")
            self._exchange = exchange
            self._shard_masses = None
            self._comm = ShardComm(group)
            self._layout = ShardLayout(n_particles, self._comm.world)
            self._base_counts = list(self._layout.counts)      # the balanced split; slabs float around it
            self._n_global = int(n_particles)
            self._mail = None
            self._ops = None
            resampler = kwargs.get('resampler')
            # (rng='numpy' / 'mt19937': the parity mode — every rank must hold the same np.random state, e.g.
            # np.random.seed(s) on every rank, because every rank generates the whole legacy stream)
            self.last_parity_js = None
            super(ShardedSMCUpdater, self).__init__(model, self._layout.count(self._comm.rank), prior, **kwargs)
            if resampler is None:
                from .resamplers import LiuWestResampler
                self.resampler = LiuWestResampler(rng='philox', scan='fast', seed=0x5EED)
            self._min_n_ess = self._n_global
            self._n_ess = float(self._n_global)
            # "split" resample: the multinomial split is drawn by every rank from the SAME counter-based host
            # generator (rank 0's resampler seed), the offspring from per-rank Philox streams
            seed0 = self._comm.broadcast_int(int(getattr(self.resampler, '_seed', 0x5EED)) & ((1 << 63) - 1),
                                             self._cloud.device)
            self._split_rng = np.random.Generator(np.random.Philox(key=seed0 & ((1 << 64) - 1)))
            self._stream_seed = (seed0 + 0x9E3779B97F4A7C15 * (self._comm.rank + 1)) & ((1 << 64) - 1)
            self.last_exchange = (0, 0)
            self._sample_calls = 0
            if _group_key(self._comm) not in _WARMED:
                self._warm_collectives()
                _WARMED.add(_group_key(self._comm))

        def _warm_collectives(self):
            """Create the NCCL channels the resample exchange uses now, not inside the first resample."""
            dev = self._cloud.device
            comm = self._comm
            comm.all_reduce_sum(torch.zeros((4,), dtype=torch.float64, device=dev))
            comm.all_gather_rows(torch.zeros((4,), dtype=torch.float64, device=dev))
            comm.all_gather_scalars(0.0, dev)
            cnt = comm.exchange_counts([1] * comm.world, dev)
            comm.all_to_all_v(torch.zeros((comm.world,), dtype=torch.float64, device=dev), [1] * comm.world, cnt, 1)
            torch.cuda.synchronize()

        # -- plumbing ----------------------------------------------------------------------
        SLAB_SLACK = 0.03       # floating slabs: capacity = balanced size * (1 + slack) (at least + 4096)

        @classmethod
        def _slab_capacity(cls, n):
            return int(n) + max(4096, int(cls.SLAB_SLACK * n))

        def _cloud_capacity(self, n):
            return self._slab_capacity(n)

        def _rebuild_cloud(self, n):
            launches_tag = self._cloud._tag if self._cloud is not None else 0
            super(ShardedSMCUpdater, self)._rebuild_cloud(n)
            if self._comm.world > 1:
                if self._mail is None:
                    self._mail = PeerMailboxes.acquire(self._comm)
                self._mail.last_tag = max(self._mail.last_tag, launches_tag)
                self._cloud._tag = self._mail.last_tag          # never reuse a tag the mailbox rows may still carry
                self._mail.install(self._cloud._ctl)
                self._cloud.enable_shard_norms()      # every update publishes the shard masses a resample splits by
            self._ops = _DeviceOps(self._cloud, self._comm.world)

        def close(self):
            """Collective: drain this updater's launches on every rank and park the peer mailboxes for the next
            sharded updater of this group (``qinfer_b200.sharded.shutdown()`` unmaps them for good)."""
            if self._mail is not None:
                self._flush()
                self._mail.release(self._cloud._tag)
                self._mail = None

        @property
        def n_particles(self):
            return self._n_global

        @property
        def n_local(self):
            return self._cloud.n

        def reset(self, n_particles=None, only_params=None, reset_weights=True):
            if self._cloud is not None and n_particles not in (None, self._cloud.n, self._n_global):
                raise ValueError("changing the particle count of a sharded cloud is not supported")
            local = self._base_counts[self._comm.rank]
            if self._cloud is not None and self._cloud.n != local:
                if only_params is not None:
                    raise ValueError("reset(only_params=...) after the slabs of a sharded cloud have floated is not "
                                     "supported")
                self._flush()
                self._cloud.resize(local)                # back to the balanced split: the prior refills every slab
                self._layout.set_counts(self._base_counts)
            super(ShardedSMCUpdater, self).reset(local if self._cloud is None else None, only_params, reset_weights)
            if reset_weights:
                self._set_global_uniform()               # w = 1 / N_global on every slab

        def _set_global_uniform(self):
            self._cloud.set_uniform_weights(self._n_global)
            self._n_ess = float(self._n_global)
            self._host_weights = None

        def _restat_global(self):
            """After the host assigned slab weights: make the stats block describe the GLOBAL weight vector."""
            cloud = self._cloud
            st = cloud.read_stats().copy()
            sums = torch.tensor([st[_lib.QB_STAT_NORM], st[_lib.QB_STAT_SUMSQ]], dtype=torch.float64,
                                device=cloud.device)
            self._comm.all_reduce_sum(sums)
            norm, sumsq = [float(v) for v in sums.tolist()]
            host = np.zeros((_lib.QB_STAT_COUNT,))
            host[_lib.QB_STAT_NORM], host[_lib.QB_STAT_SUMSQ] = norm, sumsq
            host[_lib.QB_STAT_INV_NORM] = 1.0
            host[_lib.QB_STAT_NESS] = 1.0 / sumsq
            cloud.stats.copy_(torch.from_numpy(host))
            self._n_ess = 1.0 / sumsq

        # smc.py:416-418 on a sharded cloud: the bad-weight count the kernel publishes is already global, so every rank
        # takes the clip path together; the smallest weight and the post-clip sums must be global too, or the ranks
        # would take different policy / resample decisions (and the next in-kernel all-reduce would wait for a peer
        # that never launches)
        def _pending_min_weight(self, slot):
            mine = torch.tensor([self._cloud.pending_min_weight(slot)], dtype=torch.float64, device=self._cloud.device)
            self._comm.dist.all_reduce(mine, op=self._comm.dist.ReduceOp.MIN, group=self._comm.group)
            return float(mine.item())

        def _clip_weights(self, slot):
            cloud = self._cloud
            st = np.array(cloud.clip_weights(slot), dtype=np.float64)      # local sums of the clipped slab
            sums = torch.tensor([st[_lib.QB_STAT_NORM], st[_lib.QB_STAT_SUMSQ]], dtype=torch.float64,
                                device=cloud.device)
            self._comm.all_reduce_sum(sums)
            norm, sumsq = [float(v) for v in sums.tolist()]
            st[_lib.QB_STAT_NORM], st[_lib.QB_STAT_SUMSQ] = norm, sumsq
            st[_lib.QB_STAT_INV_NORM] = 1.0
            st[_lib.QB_STAT_NESS] = 1.0 / sumsq if sumsq else np.inf
            st[_lib.QB_STAT_NBAD] = 0.0
            cloud._stats[slot].copy_(torch.from_numpy(st))
            return st

        @property
        def particle_weights(self):
            return SMCUpdater.particle_weights.fget(self)

        @particle_weights.setter
        def particle_weights(self, value):
            SMCUpdater.particle_weights.fset(self, value)
            self._restat_global()

        # -- global reductions ---------------------------------------------------------------
        @nvtx_range('qb.sharded.global_moments')
        def _global_moments(self, overlap=None):
            """Global (sum w, mean, second moment): local reduction, ONE all-gather, rows summed in rank order (so
            every rank holds identical numbers).  ``overlap``: a callable that queues device work which does not
            need the moments (the CDF pass of a resample); it runs while the collective and the host read complete."""
            cloud = self._cloud
            _lib.check(cloud.lib.qb_moments(ctypes.c_void_p(cloud.x.data_ptr()), ctypes.c_void_p(cloud.w.data_ptr()),
                                            ctypes.c_void_p(cloud.stats.data_ptr()), cloud.n, cloud.d,
                                            ctypes.c_void_p(cloud.moments_out.data_ptr()),
                                            ctypes.c_void_p(cloud.ws.data_ptr()), cloud.ws_bytes,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            cloud.launches += 2
            handle = self._comm.all_gather_rows_begin(cloud.moments_out)   # one collective, one host read
            if overlap is not None:
                overlap()
            rows = self._comm.all_gather_rows_end(handle)
            out = rows[0].copy()
            for r in range(1, rows.shape[0]):                         # fixed rank order: identical on every rank
                out += rows[r]
            self._shard_masses = rows[:, 0].copy()
            d = cloud.d
            return out[0], out[1:1 + d].copy(), out[1 + d:].reshape(d, d).copy()

        def est_mean(self):
            self._flush()
            return self._global_moments()[1]

        def est_covariance_mtx(self, corr=False):
            self._flush()
            _, mean, m2 = self._global_moments()
            cov = covariance_from_moments(mean, m2)
            if corr:
                dstd = np.sqrt(np.diag(cov))
                cov /= np.outer(dstd, dstd)
            return cov

        def sample(self, n=1):
            """distributions.py:320-333 on the sharded cloud (collective: every rank calls it and receives the SAME
            ``n`` samples).  A weighted draw of n particles = split n over the shards by their masses (one shared
            multinomial, like the resample), draw inside each slab, gather the rows."""
            self._flush()
            cloud, comm = self._cloud, self._comm
            masses = comm.all_gather_scalars(self._local_mass(), cloud.device)
            m = split_counts(self._split_rng, int(n), masses)
            mine = np.empty((0, cloud.d))
            if m[comm.rank] > 0:
                cloud._resample_scratch(cloud.n if cloud._js is None else cloud._js.numel())
                cloud.cdf(_lib.QB_SCAN_FAST)
                total = float(cloud._cdf[-1].item())
                rng = np.random.Generator(np.random.Philox(key=(self._stream_seed + self._sample_calls) & ((1 << 64) - 1)))
                u = torch.from_numpy(rng.random(m[comm.rank]) * total).to(cloud.device)
                js = torch.empty((m[comm.rank],), dtype=torch.int64, device=cloud.device)
                _lib.check(cloud.lib.qb_draw(ctypes.c_void_p(cloud._cdf.data_ptr()), cloud.n,
                                             ctypes.c_void_p(u.data_ptr()), m[comm.rank],
                                             ctypes.c_void_p(js.data_ptr()),
                                             ctypes.c_void_p(cloud.counter[1:].data_ptr()),
                                             ctypes.c_void_p(cloud.ws.data_ptr()), cloud.ws_bytes,
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
                cloud.launches += 2
                mine = cloud.x.index_select(0, js).cpu().numpy()
            self._sample_calls += 1
            parts = comm.all_gather_object(mine)
            return np.concatenate([p_ for p_ in parts if p_.shape[0]], axis=0)

        def _local_mass(self):
            """Sum of this slab's globally normalised weights."""
            cloud = self._cloud
            _lib.check(cloud.lib.qb_moments(ctypes.c_void_p(cloud.x.data_ptr()), ctypes.c_void_p(cloud.w.data_ptr()),
                                            ctypes.c_void_p(cloud.stats.data_ptr()), cloud.n, cloud.d,
                                            ctypes.c_void_p(cloud.moments_out.data_ptr()),
                                            ctypes.c_void_p(cloud.ws.data_ptr()), cloud.ws_bytes,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            cloud.launches += 2
            return float(cloud.moments_out[0].item())

        def hypothetical_update(self, outcomes, expparams, return_likelihood=False, return_normalization=False):
            """smc.py:324-386 on the sharded cloud (collective): this rank's slab of the hypothetical weights,
            normalised by the GLOBAL sums (one all-reduce of the n_outcomes x n_expparams normalisations)."""
            self._flush()
            if not isinstance(outcomes, np.ndarray):
                outcomes = np.array([outcomes])
            expparams = np.atleast_1d(expparams)
            self._count_calls(outcomes.shape[0] * self._cloud.n * expparams.shape[0])
            w_loc, L, norm_loc = self._cloud.hypothetical_update(outcomes, expparams, return_likelihood)
            eps_ = np.spacing(1)
            div_loc = np.where(np.abs(norm_loc) < eps_, 1.0, norm_loc)           # what the kernel divided by
            sums = torch.from_numpy(np.ascontiguousarray(norm_loc.reshape(-1))).to(self._cloud.device)
            self._comm.all_reduce_sum(sums)
            norm = sums.cpu().numpy().reshape(norm_loc.shape)
            div = np.where(np.abs(norm) < eps_, 1.0, norm)                       # smc.py:369-370 on the global sums
            weights = w_loc * (div_loc / div)
            out = (weights,)
            if return_likelihood:
                out += (L,)
            if return_normalization:
                out += (norm,)
            return out[0] if len(out) == 1 else out

        # -- resampling -------------------------------------------------------------------------
        @nvtx_range('qb.sharded.resample')
        def resample(self):
            self._flush()
            if self._just_resampled:
                warnings.warn("Resampling without additional data; this may not perform as desired.",
                              ResamplerWarning)
            self._just_resampled = True
            self._resample_count += 1
            ev = None
            if self._cloud.resample_events is not None:  # bench instrumentation
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            res = self.resampler
            cloud = self._cloud
            comm = self._comm
            n_local, d = cloud.n, cloud.d

            if getattr(res, '_rng', 'philox') != 'philox':
                self._parity_pass()
                self._finish_resample(ev)
                return
            split = d <= 4 and getattr(res, '_fused', False) and self._exchange == 'split'
            if split and getattr(res, '_draw', 'auto') in ('auto', 'binned') and cloud.binned_supported(n_local):
                floated = self._split_pass_binned()
                self._finish_resample(ev, weights_fused=floated)
                return
            _, mean, m2 = self._global_moments(
                overlap=(lambda: cloud.cdf(_lib.QB_SCAN_FAST_GUIDE_SCALED)) if split else None)
            cov = covariance_from_moments(mean, m2)
            a, h = res._a, res._h
            if scipy.linalg.norm(cov, 'fro') == 0:
                warnings.warn("Covariance has zero norm; adding in small covariance in resampler. "
                              "Consider increasing n_particles to improve covariance estimates.", ResamplerWarning)
                cov = res._zero_cov_comp * np.eye(cov.shape[0])
            S, S_err = sqrtm_psd(cov)
            if not np.isfinite(S_err):
                raise ResamplerError("Infinite error in computing the square root of the covariance matrix. "
                                     "Check that n_ess is not too small.")
            S = np.real(h * S)

            if split:
                self._split_pass(mean, S, a)
                self._finish_resample(ev)
                return

            cloud._resample_scratch(n_local)
            cdf = cloud.cdf(_lib.QB_SCAN_FAST)                       # local scan of globally normalised weights
            shard_total = float(cdf[-1].item())
            # counter-based streams indexed by GLOBAL slot: rank r uses the sub-stream of its slab
            goff = self._layout.offsets[comm.rank]
            base = res._philox_offset
            cloud.rng_uniform(cloud._u, n_local, res._seed, base + (goff + 1) // 2)
            rows, perm, _ = route_resample(comm, self._ops, cloud._u[:n_local], shard_total, d)
            res._philox_offset = base + (self._n_global + 1) // 2 + 1

            rows2d = rows.reshape(-1, d)
            n_iters, local_invalid, first = 0, n_local, True
            while True:
                n_iters += 1
                k = n_local if first else local_invalid
                if k > 0:
                    eps = cloud._eps[:d * k]
                    cloud.rng_normal(eps, d * k, res._seed ^ 0x9E3779B97F4A7C15,
                                     res._philox_offset + (goff * d + 1) // 2)
                    if first:
                        cloud.lw_move(mean, S, a, eps, n_local, res._postselect, x_src=rows2d, js=perm)
                    else:
                        if n_iters == 2:
                            cloud._js.copy_(perm)                     # the retry kernel indexes the same rows
                        cloud.compact_invalid(n_local)
                        cloud.lw_retry(mean, S, a, eps, k, x_src=rows2d, own_mean=True)
                first = False
                res._philox_offset += (self._n_global * d + 1) // 2 + 1   # every rank advances in lockstep
                # one collective + one host read gives every rank its own and the global invalid count
                mine = cloud.counter[:1] if k > 0 else torch.zeros((1,), dtype=torch.int64, device=cloud.device)
                allc = torch.empty((comm.world,), dtype=torch.int64, device=cloud.device)
                comm.dist.all_gather_into_tensor(allc, mine, group=comm.group)
                counts = allc.tolist()
                local_invalid = int(counts[comm.rank])
                tot = torch.tensor([float(sum(counts))])
                if tot.item() == 0:
                    break
                if n_iters >= res._maxiter:
                    warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                                   "iterations.").format(int(tot.item()), res._maxiter), ResamplerWarning)
                    break
            res.last_n_iters = n_iters

            self._finish_resample(ev)

        @nvtx_range('qb.sharded.parity_pass')
        def _parity_pass(self):
            """Parity mode: global moments, then ``parity_resample`` (chained exact scan, shared legacy stream)."""
            res, cloud = self.resampler, self._cloud
            if self._layout.counts != self._base_counts:
                raise _lib.QbError("parity-mode resample on floated slabs (mix of resamplers on one sharded cloud)")
            _, mean, m2 = self._global_moments()
            S = self._liu_west_consts(mean, m2)
            ops = _ParityOps(cloud, res, mean, S)
            n_iters, bad, js = parity_resample(self._comm, ops, self._layout, cloud.d, res._postselect, res._maxiter)
            if bad:
                warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                               "iterations.").format(bad, res._maxiter), ResamplerWarning)
            res.last_n_iters = n_iters
            self.last_parity_js = js                 # global parent index of every slot of this slab (diagnostics)
            self.last_exchange = (0, 0)

        def _split_pass(self, mean, S, a):
            """Offspring counts by a shared multinomial split, offspring drawn locally by the fused kernel, surplus
            rows moved in one all-to-all (module docstring)."""
            res, cloud, comm = self.resampler, self._cloud, self._comm
            m = split_counts(self._split_rng, self._n_global, self._shard_masses)
            cloud.preallocate_resample_slab()
            state = {'cdf': False, 'iters': 0}            # the CDF (+ guide) was queued behind the moments
            updater = self

            class Ops(object):
                def slab(self):
                    return cloud.x_alt

                def alloc(self, rows):
                    return torch.empty((rows, cloud.d), dtype=torch.float64, device=cloud.device)

                def draw_into(self, dst):
                    # ('auto' stays with the guided draw here: the merge draw on a scaled slab CDF is selected
                    # only when asked for explicitly)
                    it, bad = res._fused_pass(cloud, mean, S, a, dst.shape[0], dst=dst, scale_u=True, own_mean=True,
                                              seed=updater._stream_seed, build_cdf=state['cdf'], auto_merge=False)
                    state['cdf'] = False
                    state['iters'] = max(state['iters'], it)
                    if bad:
                        warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                                       "iterations.").format(bad, res._maxiter), ResamplerWarning)

            self.last_exchange = split_resample(comm, Ops(), m, self._layout.counts, cloud.d)
            res.last_n_iters = state['iters']

        def _liu_west_consts(self, mean, m2):
            """cov, zero-norm replacement and S = h * sqrtm_psd(cov) on the host (resamplers.py:266-305), identical on
            every rank because the moments are."""
            res = self.resampler
            cov = _cov_1x1(mean, m2) if mean.shape[0] == 1 else covariance_from_moments(mean, m2)
            if (cov[0, 0] == 0) if cov.shape == (1, 1) else (scipy.linalg.norm(cov, 'fro') == 0):
                warnings.warn("Covariance has zero norm; adding in small covariance in resampler. "
                              "Consider increasing n_particles to improve covariance estimates.", ResamplerWarning)
                cov = res._zero_cov_comp * np.eye(cov.shape[0])
            S, S_err = sqrtm_psd(cov)
            if not np.isfinite(S_err):
                raise ResamplerError("Infinite error in computing the square root of the covariance matrix. "
                                     "Check that n_ess is not too small.")
            return np.real(res._h * S)

        @nvtx_range('qb.sharded.split_pass_binned')
        def _split_pass_binned(self):
            """The "split" resample over the binned draw (csrc/qb_binned.cu): ONE pass over the slab gives its bin
            sums and its moment sums; ONE all-gather of the 1 + d + d^2 sums gives the global mean / covariance and
            the shard masses; every rank draws the same multinomial split m ~ Mult(N, masses), then draws ITS m[r]
            offspring from its own slab — the first min(m[r], N/G) straight into its new slab, the surplus into the
            send buffer of the one all-to-all — in one launch."""
            res, cloud, comm = self.resampler, self._cloud, self._comm
            r, d = comm.rank, cloud.d
            cloud._binned_scratch(cloud.capacity)                # (re-allocation zeroes the workspace: before pass 1)
            masses = cloud.shard_masses(comm.world) if comm.world > 1 else None
            if masses is not None:
                cloud.binned_sums(mirror=False)                  # (pass 1 needs nothing from the host: queue it first)
                if self._queued_split_pass(masses):
                    return True
            cloud.binned_sums()
            rows = comm.all_gather_rows(cloud.moments_out)       # one collective, one host read
            out = rows[0].copy()
            for q in range(1, rows.shape[0]):                    # fixed rank order: identical on every rank
                out += rows[q]
            self._shard_masses = rows[:, 0].copy()
            mean, m2 = out[1:1 + d].copy(), out[1 + d:].reshape(d, d).copy()
            m = split_counts(self._split_rng, self._n_global, self._shard_masses)
            caps = [self._slab_capacity(c) for c in self._base_counts]
            if all(1 <= m[q] <= caps[q] for q in range(comm.world)):
                # FLOATING SLABS: every rank keeps exactly the offspring it drew — no row leaves its GPU, and the new
                # weights 1/N are written by the same launch.  Slab sizes then drift around the balanced split by
                # ~sqrt(N/G) per resample; the exchange below runs only when one outgrows its capacity.
                cloud._alt_slab(m[r])
                off_u, off_v = res._binned_offsets(m[r])
                cloud.binned_count(m[r], self._stream_seed, off_u)   # (runs while the host takes the square root)
                S = self._liu_west_consts(mean, m2)
                iters, bad = res._binned_move(cloud, mean, S, res._a, m[r], off_v, seed=self._stream_seed,
                                              fuse_weights=True, n_global=self._n_global, dst=cloud.x_alt)
                if bad:
                    warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                                   "iterations.").format(bad, res._maxiter), ResamplerWarning)
                self._layout.set_counts(m)
                self.last_exchange = (0, 0)
                res.last_n_iters = iters
                return True
            S = self._liu_west_consts(mean, m2)
            cap = self._base_counts                              # rebalance: everybody back to the balanced split
            cloud._alt_slab(cap[r])
            keep = min(m[r], cap[r])
            extra = m[r] - keep
            send = torch.empty((max(extra, 1), d), dtype=torch.float64, device=cloud.device)
            iters = 0
            if m[r] > 0:
                if m[r] > cloud._bin_cap[1]:                     # a slab holding far more than its share of the mass
                    cloud._binned_scratch(m[r])
                    cloud.binned_sums()
                off_u, off_v = res._binned_offsets(m[r])
                cloud.binned_count(m[r], self._stream_seed, off_u)
                iters, bad = res._binned_move(cloud, mean, S, res._a, m[r], off_v, seed=self._stream_seed,
                                              dst=cloud.x_alt, split=keep, dst2=send if extra else None)
                if bad:
                    warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                                   "iterations.").format(bad, res._maxiter), ResamplerWarning)
            T = exchange_plan(m, cap)
            if any(any(row) for row in T):
                recv_counts = [T[q][r] for q in range(comm.world)]
                assert sum(recv_counts) == cap[r] - keep
                comm.all_to_all_into(cloud.x_alt[keep:].reshape(-1), send[:extra], T[r], recv_counts, d)
            self._layout.set_counts(cap)
            self.last_exchange = (extra, cap[r] - keep)
            res.last_n_iters = iters
            return False

        def _queued_split_pass(self, masses):
            """The floating-slab resample with NO host round trip between its launches.  The shard masses came with the
            last update's in-kernel all-reduce (every rank's own sum w', identical on every rank), so the split m is
            drawn BEFORE anything is queued; then pass 1, the all-gather of the moment sums (device to device), the
            kernel that sums them in rank order and derives the Liu-West constants, the counts, the move and the first
            retry launch go out back to back (pass 1 has been queued by the caller) and the host waits once — its checks (finite / PSD / zero-norm warnings)
            read the published global moments while the device is already drawing.  False: the split does not fit
            the slabs' capacities (the caller then runs the exchange path; the shared generator has advanced on every
            rank alike)."""
            res, cloud, comm = self.resampler, self._cloud, self._comm
            r, d = comm.rank, cloud.d
            m = split_counts(self._split_rng, self._n_global, masses)
            caps = [self._slab_capacity(c) for c in self._base_counts]
            if not all(1 <= m[q] <= caps[q] for q in range(comm.world)):
                return False
            cloud._alt_slab(m[r])
            width = 1 + d + d * d
            rows = getattr(self, '_moment_rows', None)
            if rows is None or rows.numel() != comm.world * width:
                rows = self._moment_rows = torch.empty((comm.world * width,), dtype=torch.float64, device=cloud.device)
            comm.dist.all_gather_into_tensor(rows, cloud.moments_out[:width], group=comm.group)
            tag = cloud.binned_shard_consts(rows, comm.world, res._a, res._h, res._zero_cov_comp)
            off_u, off_v = res._binned_offsets(m[r])
            cloud.binned_count(m[r], self._stream_seed, off_u)
            seed = self._stream_seed
            seed_n = seed ^ 0x9E3779B97F4A7C15
            stride, off_n, rounds = res._binned_plan(d, m[r])
            tags = cloud.binned_move(None, None, res._a, seed, off_v, seed_n, off_n, m[r], res._postselect,
                                     dst=cloud.x_alt, fuse_weights=True, n_global=self._n_global,
                                     retry_rounds=rounds, own_mean=res._own_mean)
            mtag = tags[0] if rounds else tags
            # the reference's checks on the global moments (distributions.py:388-397, resamplers.py:288-299)
            _, mean, m2 = cloud.binned_moments_wait(tag)
            flag, s_err = cloud.binned_flags()
            _cov_1x1(mean, m2) if d == 1 else covariance_from_moments(mean, m2)
            if flag == 1:
                warnings.warn("Covariance has zero norm; adding in small covariance in resampler. "
                              "Consider increasing n_particles to improve covariance estimates.", ResamplerWarning)
            if not np.isfinite(s_err):
                raise ResamplerError("Infinite error in computing the square root of the covariance matrix. "
                                     "Check that n_ess is not too small.")
            iters, bad = res._binned_finish(cloud, mtag, rounds, stride, None, None, res._a, m[r], seed_n,
                                            dst=cloud.x_alt, seed_v=seed)
            if bad:
                warnings.warn(("Liu-West resampling failed to find valid models for {} particles within {} "
                               "iterations.").format(bad, res._maxiter), ResamplerWarning)
            self._shard_masses = np.asarray(masses, dtype=np.float64)
            self._layout.set_counts(m)
            self.last_exchange = (0, 0)
            res.last_n_iters = iters
            return True

        def _finish_resample(self, ev, weights_fused=False):
            cloud = self._cloud
            n_new = self._layout.count(self._comm.rank)
            if weights_fused:                                         # the move wrote 1/N and the global stats block
                cloud.adopt_binned(n_new, True)
                self._n_ess = float(self._n_global)
                self._host_weights = None
            else:
                cloud._swap_slabs(n_new)
                self._set_global_uniform()                            # weights buffer stays, contents reset
            self._host_locs = self._host_weights = None
            if self._canonicalize:
                cloud.canonicalize()
            try:
                self.model.clear_cache()
            except Exception as e:  # pragma: no cover
                warnings.warn("Exception raised when clearing model cache: {}. Ignoring.".format(e))
            if ev is not None:
                ev[1].record()
                self._cloud.resample_events.append(ev)

    return ShardedSMCUpdater


def shutdown():
    """Unmap and free every cached peer mailbox (collective; call before destroying the process group if the process
    goes on living — at interpreter exit the driver does it)."""
    for key in list(_MAILBOX_CACHE):
        _MAILBOX_CACHE.pop(key).close()
    _WARMED.clear()


_cls = None


def __getattr__(name):
    global _cls
    if name == 'ShardedSMCUpdater':
        if _cls is None:
            _cls = _make_sharded_updater_class()
        return _cls
    raise AttributeError(name)
