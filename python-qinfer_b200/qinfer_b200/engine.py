"""Device-resident particle cloud: torch tensors for memory/streams, the C ABI for all math.

``DeviceCloud`` is the single owner of the HBM state of one updater (or one
shard of it): the (n, d) fp64 particle slab (double-buffered for resampling),
the unnormalised weight vector (ping-pong so a rejected update leaves the old
weights intact, smc.py:423-441), the 16-double stats block and the kernel
workspaces.  Every arithmetic step is a call into ``libqinfer_b200.so``;
nothing here computes on the CPU and nothing falls back to it.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import nvtx_range
import time

from ._lib import (QB_MAX_FUSE, QB_STAT_COUNT, QB_STAT_INV_NORM, QB_STAT_MIN, QB_STAT_NBAD, QB_STAT_NESS,
                   QB_STAT_NORM, QB_STAT_SKIPPED, QB_STAT_SUMSQ, QB_STAT_TAG, check)

_U64_MASK = (1 << 64) - 1
MIRROR_SLOT = 8 * QB_MAX_FUSE      # doubles per mirror slot: one 8-double block per fused step


def _require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.QbError("qinfer_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    return dev


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


SHARD_NORMS_SLOT = 4 * ((_lib.QB_MAX_RANKS + 2) // 3)     # doubles per slot of qb_update_ctl.h_shard_norms


class DeviceCloud(object):
    def __init__(self, desc, n, device=None, capacity=None):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.desc = desc
        self.n = int(n)
        self.d = int(desc.d)
        # ``capacity`` >= n particles are allocated; the cloud may then change its size up to that bound without
        # touching the allocator (a resample into another particle count; the floating slabs of a sharded cloud)
        self.capacity = max(self.n, int(capacity or 0))
        with torch.cuda.device(self.device):
            f64 = dict(dtype=torch.float64, device=self.device)
            self._x_back = torch.empty((self.capacity, self.d), **f64)
            self._x_alt_back = None               # allocated at the first resample
            self.x = self._x_back[:self.n]
            self.x_alt = None
            _pad = int(os.environ.get("QB_ALLOC_PAD", "0"))     # experiment knob: shift the weight buffers
            self._pad_tensor = torch.empty((_pad,), dtype=torch.uint8, device=self.device) if _pad else None
            # weights and stats are ping-pong pairs: `cur` is the committed state, 1 - cur receives the
            # next update (so a rejected update leaves the committed weights intact, smc.py:423-441)
            self._w_back = [torch.empty((self.capacity,), **f64), torch.empty((self.capacity,), **f64)]
            self._w = [b[:self.n] for b in self._w_back]
            self._stats = [torch.zeros((QB_STAT_COUNT,), **f64), torch.zeros((QB_STAT_COUNT,), **f64)]
            self.cur = 0
            # host-visible copy of each stats block, written by the kernel itself (pinned, device-accessible)
            self.mirror = torch.zeros((2 * MIRROR_SLOT,), dtype=torch.float64, pin_memory=True)
            self.mirror_np = self.mirror.numpy()
            self._ctl = _lib.QbUpdateCtl()
            self._tag = 0
            self._eps_arr = (_lib.QbExpparams * QB_MAX_FUSE)()
            self._out_arr = (ctypes.c_int64 * QB_MAX_FUSE)()
            self._alloc_workspace()
            self.moments_out = torch.empty((1 + self.d + self.d * self.d,), **f64)
            self.moments_host = torch.empty((1 + self.d + self.d * self.d,), dtype=torch.float64, pin_memory=True)
            self.stats_host = torch.empty((QB_STAT_COUNT,), dtype=torch.float64, pin_memory=True)
            self.counter = torch.zeros((4,), dtype=torch.int64, device=self.device)
            self.counter_host = torch.empty((2,), dtype=torch.int64, pin_memory=True)
            self.basis_dev = None
            if desc.basis is not None:
                b = np.ascontiguousarray(desc.basis).view(np.float64).reshape(-1)
                self.basis_dev = torch.from_numpy(b.copy()).to(self.device)
            self.lib_model = ctypes.pointer(desc.c_model)
        # resample scratch, allocated lazily
        self._cdf = self._js = self._u = self._eps = self._invalid = self._idxs = self._parent_inv = None
        self._moments_event = None
        # binned resample (qb_lw_binned_*): private zeroed workspace, invalid-slot list and a pinned mirror the
        # kernels write their results into (moments: 32 doubles; counters: 4 doubles), allocated at first use
        self._bin_ws = None
        self._bin_cap = (0, 0)
        self._bin_list = None
        self._bin_mirror = None
        self._bin_tag = 0
        self.count_mode = int(os.environ.get('QB_BINNED_COUNT', _lib.QB_COUNT_AUTO))   # pass 2: auto | histogram | binomial tree
        # noise source of the f4 decorators (PoisonedModel, random walks): see ``normals``
        self.noise_rng, self.noise_seed, self.noise_offset = 'numpy', 0x6E6F697365, 0
        # one control block per destination slot: the constant fields are written once
        self._ctls = [_lib.QbUpdateCtl(), _lib.QbUpdateCtl()]
        self._ctl_key = None
        self._shard_norms = None           # sharded clouds: pinned [2 slots][QB_MAX_RANKS + 1] per-rank sums + tag
        self._slot_tag = [0, 0]            # tag of the update launch whose output weights buffer [slot] still holds
        self._refresh_ptrs()
        self._chain_tag = 0                # tag of the fused update that was the LAST thing queued for this cloud
        self._chain_dst = -1               # ... and the weights/stats buffer it wrote
        self._launches = 0
        self.resample_events = None        # bench: set to [] to collect a CUDA-event pair around every resample
        self.update_launches = 0

    def _alloc_workspace(self):
        cap = self.capacity
        ws_bytes = max(self.lib.qb_update_workspace_bytes(cap, self.d),
                       self.lib.qb_moments_workspace_bytes(cap, self.d),
                       self.lib.qb_cdf_workspace_bytes(cap),
                       self.lib.qb_compact_workspace_bytes(cap),
                       self.lib.qb_draw_workspace_bytes(cap))
        self.ws = torch.zeros(((ws_bytes + 7) // 8,), dtype=torch.float64, device=self.device)   # zeroed once: holds the launch ticket
        self.ws_bytes = self.ws.numel() * 8

    def _alt_slab(self, n_new):
        """The alternate particle slab as an (n_new, d) view of its backing buffer (grown only when too small)."""
        n_new = int(n_new)
        if self._x_alt_back is None or self._x_alt_back.shape[0] < n_new:
            self._x_alt_back = torch.empty((max(n_new, self.capacity), self.d), dtype=torch.float64,
                                           device=self.device)
            self.x_alt = None
        if self.x_alt is None or self.x_alt.shape[0] != n_new:
            self.x_alt = self._x_alt_back[:n_new]
        return self.x_alt

    def _refresh_ptrs(self):
        """Raw pointers of the buffers the hot loop passes on every launch (re-derived whenever a buffer is swapped)."""
        self._px = _ptr(self.x)
        self._pw = [_ptr(self._w[0]), _ptr(self._w[1])]
        self._pstats = [_ptr(self._stats[0]), _ptr(self._stats[1])]
        self._pws = _ptr(self.ws)
        self._px_id = self.x.data_ptr()

    # Every entry point other than the fused update bumps ``launches`` after queueing its kernels; that also breaks
    # the update chain (a chained update depends on its predecessor through flags, not on whatever ran in between).
    @property
    def launches(self):
        return self._launches

    @launches.setter
    def launches(self, value):
        self._launches = value
        self._chain_tag = 0

    # committed / pending views of the ping-pong buffers
    w = property(lambda self: self._w[self.cur])
    w_alt = property(lambda self: self._w[1 - self.cur])
    stats = property(lambda self: self._stats[self.cur])
    stats_alt = property(lambda self: self._stats[1 - self.cur])

    # ---- host <-> device ---------------------------------------------------
    def upload_locations(self, locs):
        locs = np.ascontiguousarray(locs, dtype=np.float64)
        if locs.shape != (self.n, self.d):
            raise ValueError("particle_locations must have shape (%d, %d), got %s" % (self.n, self.d, locs.shape))
        self.x.copy_(torch.from_numpy(locs))
        self._chain_tag = 0

    def _to_host(self, t):
        """Device -> host through a pinned staging tensor.  torch's caching host allocator recycles pinned blocks, so
        in a long-lived process the read-back runs at PCIe speed; only the very first read of a given size pays the
        page-locking (cudaHostAlloc, ~0.6 ms/MB measured).  The returned NumPy array owns its (pinned) memory."""
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h.numpy()

    def download_locations(self):
        return self._to_host(self.x)

    def upload_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        if w.shape != (self.n,):
            raise ValueError("particle_weights must have shape (%d,), got %s" % (self.n, w.shape))
        self.w.copy_(torch.from_numpy(w))
        self._slot_tag = [0, 0]
        check(self.lib.qb_weights_restat(_ptr(self.w), self.n, _ptr(self.stats), _ptr(self.ws), self.ws_bytes,
                                         _stream()))
        self.launches += 2

    def download_weights(self):
        """Normalised weights, materialised by a kernel then copied out."""
        out = self.w_alt
        check(self.lib.qb_weights_normalized(_ptr(self.w), self.n, _ptr(self.stats), _ptr(out), _stream()))
        self.launches += 1
        return self._to_host(out)

    def set_uniform_weights(self, n_global=None):
        n_global = self.n if n_global is None else int(n_global)
        check(self.lib.qb_weights_set_uniform_global(_ptr(self.w), self.n, n_global, _ptr(self.stats), _stream()))
        self.launches += 1
        self._slot_tag = [0, 0]

    def read_stats(self, which=None):
        """Blocking read of a stats block (norm, sumsq, min, nbad, inv_norm, n_ess)."""
        src = self.stats if which is None else which
        self.stats_host.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.stats_host.numpy()

    # ---- the hot kernel -------------------------------------------------------
    @nvtx_range('qb.cloud.fused_update')
    def fused_update(self, steps, src, guard=False, zero_weight_thresh=0.0, resample_below=0.0):
        """Launch ONE fused kernel applying the consecutive updates ``steps`` = [(ep_record, outcome, check), ...]
        (1..QB_MAX_FUSE of them) to weights/stats buffer ``src``, writing buffer ``1 - src``; the kernel mirrors one
        8-double block per step into pinned host slot ``1 - src``.  Returns the launch tag.  ``guard``: speculative
        launch that cancels itself if the launch that produced ``src`` needs the host."""
        dst = 1 - src
        k = len(steps)
        self._tag += 1
        if self.x.data_ptr() != self._px_id or self._pw[0].value != self._w[0].data_ptr():
            self._refresh_ptrs()                       # a resample swapped the slabs (or re-sized the weights)
        key = (zero_weight_thresh, resample_below, self._ctl.n_ranks)
        if key != self._ctl_key:                       # constants of the two per-slot control blocks
            for slot, c in enumerate(self._ctls):
                ctypes.memmove(ctypes.byref(c), ctypes.byref(self._ctl), ctypes.sizeof(_lib.QbUpdateCtl))
                c.h_mirror = self.mirror.data_ptr() + slot * MIRROR_SLOT * 8
                c.h_shard_norms = (self._shard_norms.data_ptr() + slot * SHARD_NORMS_SLOT * 8
                                   if self._shard_norms is not None else None)
                c.zero_weight_thresh = zero_weight_thresh
                c.resample_below = resample_below
            self._ctl_key = key
        ctl = self._ctls[dst]
        ctl.tag = float(self._tag)
        ctl.guard = 1 if guard else 0
        # chained launch: the previous thing queued for this cloud is the update whose output this one reads
        ctl.chain_prev_tag = float(self._chain_tag) if (self._chain_tag and self._chain_dst == src) else 0.0
        stream = _stream()
        if k == 1 and self.desc.poison is not None:
            ep, outcome, chk = steps[0]
            ctl.check_resample = 1 if chk else 0
            self._poisoned_update(ep, outcome, src, dst, ctl, stream)
        elif k == 1:
            ep, outcome, chk = steps[0]
            ctl.check_resample = 1 if chk else 0
            check(self.lib.qb_fused_update(self.lib_model, ctypes.byref(ep), int(outcome), self._px, self.n,
                                           self._pw[src], self._pw[dst], self._pstats[src], self._pstats[dst],
                                           ctypes.byref(ctl), self._pws, self.ws_bytes, stream))
        else:
            mask = 0
            eps, outs = self._eps_arr, self._out_arr
            for j, (ep, outcome, chk) in enumerate(steps):
                ctypes.memmove(ctypes.byref(eps, j * ctypes.sizeof(_lib.QbExpparams)), ctypes.byref(ep),
                               ctypes.sizeof(_lib.QbExpparams))
                outs[j] = int(outcome)
                if chk:
                    mask |= 1 << j
            check(self.lib.qb_fused_update_multi(self.lib_model, eps, outs, k, mask, self._px, self.n,
                                                 self._pw[src], self._pw[dst], self._pstats[src], self._pstats[dst],
                                                 None, ctypes.byref(ctl), self._pws, self.ws_bytes, stream))
        self.launches += 1
        self.update_launches += 1
        self._chain_tag, self._chain_dst = self._tag, dst
        self._slot_tag[dst] = self._tag
        return self._tag

    def enable_shard_norms(self):
        """Sharded clouds: every update launch also publishes each rank's own sum w' (the shard masses)."""
        if self._shard_norms is None:
            self._shard_norms = torch.zeros((2 * SHARD_NORMS_SLOT,), dtype=torch.float64, pin_memory=True)
            self._shard_norms_np = self._shard_norms.numpy()
            self._ctl_key = None

    def shard_masses(self, n_ranks):
        """Per-rank sums of the committed weights buffer as the update launch that wrote it published them (identical
        on every rank), or None when the buffer has been written by anything else since."""
        tag = self._slot_tag[self.cur]
        if self._shard_norms is None or not tag:
            return None
        row = self._shard_norms_np[self.cur * SHARD_NORMS_SLOT:(self.cur + 1) * SHARD_NORMS_SLOT].reshape(-1, 4)
        groups = (n_ranks + 2) // 3
        # groups of {3 sums, tag}, one 32-byte store each, issued right after the stats block the host has already
        # seen: wait for them (whether the masses exist must not depend on timing — every rank takes the same path)
        spins, t0 = 0, None
        while not np.all(row[:groups, 3] == float(tag)):
            spins += 1
            if spins > 2000:
                if t0 is None:
                    t0 = time.perf_counter()
                elif time.perf_counter() - t0 > 10.0:
                    raise _lib.QbError("timed out waiting for the shard masses of update launch %d" % tag)
        return row[:groups, :3].reshape(-1)[:n_ranks].copy()

    @nvtx_range('qb.cloud.wait_stats')
    def wait_stats(self, slot, tag, nsteps=1, timeout_s=120.0):
        """Spin on the pinned mirror of stats buffer ``slot`` until launch ``tag`` has published its ``nsteps`` blocks.
        Each block is {S, Q, #bad, TAG | record, n_ess, TAG, attention + 2*skipped}, written as two 32-byte stores;
        it is accepted when both tags match.  Returns an (nsteps, 8) array."""
        m = self.mirror_np
        base = slot * MIRROR_SLOT
        want = float(tag)
        if nsteps == 1:                         # the common case: two scalar reads per poll
            i3, i6 = base + 3, base + 6
            spins = 0
            while True:
                if m[i3] == want and m[i6] == want:
                    a = m[base:base + 8].copy()
                    if a[3] == want and a[6] == want:
                        return a.reshape(1, 8)
                spins += 1
                if spins > 20000:
                    break                       # fall through to the timed loop below
        t0 = None
        last = base + 8 * (nsteps - 1)
        while True:
            a = m[base:base + 8 * nsteps].copy()
            if a[3] == want and a[6] == want and a[last - base + 3] == want and a[last - base + 6] == want:
                a = a.reshape(nsteps, 8)
                if np.all(a[:, 3] == want) and np.all(a[:, 6] == want):
                    return a
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout_s:
                torch.cuda.synchronize()           # surfaces a CUDA error if the kernel died
                raise _lib.QbError("timed out waiting for the fused-update kernel (tag %d)" % tag)

    def pending_min_weight(self, slot):
        """Smallest weight of weights buffer ``slot``, for the warning text of smc.py:417."""
        out = self.moments_out[:1]
        check(self.lib.qb_weights_min(_ptr(self._w[slot]), self.n, _ptr(out), _stream()))
        self.launches += 1
        return float(out.cpu().numpy()[0])

    def commit_update(self):
        self.cur = 1 - self.cur

    def clip_weights(self, slot):
        """smc.py:416-418 on weights buffer ``slot`` (normalise, clip to [0,1], re-derive its stats)."""
        check(self.lib.qb_weights_clip(_ptr(self._w[slot]), self.n, _ptr(self._stats[slot]), _ptr(self.ws),
                                       self.ws_bytes, _stream()))
        self._slot_tag = [0, 0]
        self.launches += 2
        return self.read_stats(self._stats[slot])

    def hypothetical_update(self, outcomes, expparams, want_likelihood):
        """smc.py:324-386 on the device: (weights (n_o, n_e, n), L or None, norms (n_o, n_e, 1)) as host arrays."""
        outcomes = np.atleast_1d(np.asarray(outcomes)).astype(np.int64)
        expparams = np.atleast_1d(np.asarray(expparams))
        n_o, n_e = outcomes.shape[0], expparams.shape[0]
        eps = (_lib.QbExpparams * n_e)(*[self.desc.expparams_record(expparams, e) for e in range(n_e)])
        outs = (ctypes.c_int64 * n_o)(*[int(o) for o in outcomes])
        f64 = dict(dtype=torch.float64, device=self.device)
        wts = torch.empty((n_o, n_e, self.n), **f64)
        L = torch.empty((n_o, n_e, self.n), **f64) if want_likelihood else None
        norms = torch.empty((n_o, n_e), **f64)
        check(self.lib.qb_hypothetical_update(self.lib_model, eps, n_e, outs, n_o, _ptr(self.x), _ptr(self.w),
                                              _ptr(self.stats), self.n, _ptr(wts), _ptr(L) if want_likelihood else None,
                                              _ptr(norms), _ptr(self.ws), self.ws_bytes, _stream()))
        self.launches += 3 * n_o * n_e
        return (wts.cpu().numpy(), L.cpu().numpy() if want_likelihood else None,
                norms.cpu().numpy()[..., np.newaxis])

    def design_sums(self, expparams, idx, outcomes, centre, want_kld):
        """qb_design_sums for experiment ``expparams[idx]``: (sums (n_o, 1 + 2d), kld (n_o,) or None) on the host."""
        outcomes = np.atleast_1d(np.asarray(outcomes)).astype(np.int64)
        n_o = outcomes.shape[0]
        ep = self.desc.expparams_record(expparams, idx)
        outs = (ctypes.c_int64 * n_o)(*[int(o) for o in outcomes])
        need = self.lib.qb_design_workspace_bytes(self.n, self.d, n_o)
        ws = getattr(self, '_design_ws', None)
        if ws is None or ws.numel() * 8 < need:
            self._design_ws = ws = torch.zeros(((need + 7) // 8,), dtype=torch.float64, device=self.device)
        out = torch.empty((n_o * (2 + 2 * self.d),), dtype=torch.float64, device=self.device)
        sums, kld = out[:n_o * (1 + 2 * self.d)], out[n_o * (1 + 2 * self.d):]
        check(self.lib.qb_design_sums(self.lib_model, ctypes.byref(ep), outs, n_o, _ptr(self.x), _ptr(self.w),
                                      _ptr(self.stats), self.n, _lib.f64_array(centre), _ptr(sums),
                                      _ptr(kld) if want_kld else None, _ptr(ws), ws.numel() * 8, _stream()))
        self.launches += 4 if want_kld else 2
        host = out.cpu().numpy()
        return (host[:n_o * (1 + 2 * self.d)].reshape(n_o, 1 + 2 * self.d).copy(),
                host[n_o * (1 + 2 * self.d):].copy() if want_kld else None)

    # ---- moments ------------------------------------------------------------------
    def moments_begin(self):
        """Launch the moment reduction and its read-back; returns at once.  Work queued after this call (e.g. the CDF
        pass of a resample, which does not need the moments) runs while the host waits in ``moments_end``."""
        check(self.lib.qb_moments(_ptr(self.x), _ptr(self.w), _ptr(self.stats), self.n, self.d,
                                  _ptr(self.moments_out), _ptr(self.ws), self.ws_bytes, _stream()))
        self.launches += 2
        self.moments_host.copy_(self.moments_out, non_blocking=True)
        if self._moments_event is None:
            self._moments_event = torch.cuda.Event()
        self._moments_event.record()

    def moments_end(self):
        self._moments_event.synchronize()
        out = self.moments_host.numpy()
        d = self.d
        return out[0], out[1:1 + d].copy(), out[1 + d:].reshape(d, d).copy()

    def moments(self):
        """(sum w, mean (d,), second moment (d, d)) of the normalised cloud."""
        self.moments_begin()
        return self.moments_end()

    # ---- resampling -----------------------------------------------------------------
    def _resample_scratch(self, n_new):
        dev = self.device
        if self._cdf is None or self._cdf.numel() != self.n:
            self._cdf = torch.empty((self.n,), dtype=torch.float64, device=dev)
        if self._js is None or self._js.numel() != n_new:
            self._js = torch.empty((n_new,), dtype=torch.int64, device=dev)
            self._u = torch.empty((n_new,), dtype=torch.float64, device=dev)
            self._eps = torch.empty((self.d * n_new,), dtype=torch.float64, device=dev)
            self._invalid = torch.zeros((n_new,), dtype=torch.uint8, device=dev)
            self._idxs = torch.empty((n_new,), dtype=torch.int64, device=dev)

    def _fused_scratch(self, cap):
        """Scratch of the fused draw+move path: only the validity flags and the compacted retry list (capacity-based)."""
        if self._invalid is None or self._invalid.numel() < cap:
            self._invalid = torch.zeros((cap,), dtype=torch.uint8, device=self.device)
            self._idxs = torch.empty((cap,), dtype=torch.int64, device=self.device)
        if self._cdf is None or self._cdf.numel() != self.n:
            self._cdf = torch.empty((self.n,), dtype=torch.float64, device=self.device)

    def preallocate_resample(self, n_new=None):
        """Allocate the resampling scratch and the second particle slab up front (no allocation in the loop)."""
        n_new = self.n if n_new is None else int(n_new)
        self._resample_scratch(n_new)
        self._alt_slab(n_new)
        if self._parent_inv is None or self._parent_inv.numel() < n_new:
            self._parent_inv = torch.empty((n_new,), dtype=torch.int32, device=self.device)

    def preallocate_resample_slab(self):
        self._alt_slab(self.n)

    def cdf(self, mode, carry=None):
        """``carry``: a one-element device tensor holding the exact running sum of the slabs before this one (the
        chained exact scan of a sharded cloud's parity mode)."""
        if self._cdf is None or self._cdf.numel() != self.n:
            self._cdf = torch.empty((self.n,), dtype=torch.float64, device=self.device)
        if carry is not None:
            if mode != _lib.QB_SCAN_EXACT:
                raise ValueError("a carried-in running sum needs the exact scan")
            check(self.lib.qb_cdf_chained(_ptr(self.w), _ptr(self.stats), self.n, _ptr(self._cdf), _ptr(carry),
                                          _ptr(self.ws), self.ws_bytes, _stream()))
            self.launches += 1
            return self._cdf
        check(self.lib.qb_cdf(_ptr(self.w), _ptr(self.stats), self.n, _ptr(self._cdf), int(mode), _ptr(self.ws),
                              self.ws_bytes, _stream()))
        self.launches += 1 if mode == _lib.QB_SCAN_EXACT else 3
        return self._cdf

    def exact_scan_fell_back(self):
        """Diagnostics: did the last exact scan hand over to the sequential kernel? (synchronises)"""
        flag = ctypes.c_int32(-1)
        check(self.lib.qb_cdf_exact_fallback_flag(_ptr(self.ws), self.n, ctypes.byref(flag), _stream()))
        return int(flag.value)

    def draw(self, u_dev, n_new):
        check(self.lib.qb_draw(_ptr(self._cdf), self.n, _ptr(u_dev), int(n_new), _ptr(self._js),
                               _ptr(self.counter[1:]), _ptr(self.ws), self.ws_bytes, _stream()))
        self.launches += 2
        return self._js

    def _lw_consts(self):
        """This cloud's device buffer for S and (1 - a) mean of the staged Liu-West kernels (caller-owned)."""
        buf = getattr(self, '_lw_consts_buf', None)
        if buf is None:
            nbytes = self.lib.qb_lw_move_workspace_bytes(self.d)
            self._lw_consts_buf = buf = torch.empty(((nbytes + 7) // 8,), dtype=torch.float64, device=self.device)
        return buf

    def lw_move(self, mean, S, a, eps_dev, n_new, postselect, x_src=None, js=None):
        """x_alt[i] = a * x_src[js[i]] + (1-a) * mean + S @ eps[:, i] (x_src defaults to the current slab)."""
        self._alt_slab(n_new)
        src = self.x if x_src is None else x_src
        js = self._js if js is None else js
        check(self.lib.qb_lw_move(self.lib_model, _ptr(src), src.shape[0], self.d, _ptr(js),
                                  _lib.f64_array(mean), _lib.f64_array(np.asarray(S).reshape(-1)), float(a),
                                  _ptr(eps_dev), int(n_new), _ptr(self.x_alt), int(bool(postselect)),
                                  _ptr(self._invalid), _ptr(self.counter), _ptr(self._lw_consts()),
                                  self._lw_consts().numel() * 8, _stream()))
        self.launches += 1

    def lw_draw_move(self, mean, S, a, seed_u, off_u, seed_n, off_n, n_new, postselect, dst=None, scale_u=False,
                     use_guide=True):
        """Fused first Liu-West pass (device RNG, d <= 4): dst[i] = a * x[draw(u_i)] + (1-a) * mean + S @ eps[:, i]
        with u and eps regenerated in the kernel from the Philox streams (seed_u, off_u) / (seed_n, off_n).
        ``dst`` defaults to the alternate slab."""
        if dst is None:
            self._alt_slab(n_new)
            dst = self.x_alt
        self._fused_scratch(n_new)
        check(self.lib.qb_lw_draw_move(self.lib_model, _ptr(self.x), self.n, self.d, _ptr(self._cdf), _ptr(self.ws),
                                       self.ws_bytes, 1 if use_guide else 0, _lib.f64_array(mean),
                                       _lib.f64_array(np.asarray(S).reshape(-1)), float(a),
                                       int(seed_u) & _U64_MASK, int(off_u) & _U64_MASK, int(seed_n) & _U64_MASK,
                                       int(off_n) & _U64_MASK, 1 if scale_u else 0, int(n_new), _ptr(dst),
                                       int(bool(postselect)), _ptr(self._invalid), _ptr(self.counter), _stream()))
        self.launches += 1

    def lw_draw_retry(self, mean, S, a, seed_u, off_u, seed_n, off_n, k, dst=None, scale_u=False, use_guide=True,
                      own_mean=False):
        dst = self.x_alt if dst is None else dst
        check(self.lib.qb_lw_draw_retry(self.lib_model, _ptr(self.x), self.n, self.d, _ptr(self._cdf), _ptr(self.ws),
                                        self.ws_bytes, 1 if use_guide else 0, _lib.f64_array(mean),
                                        _lib.f64_array(np.asarray(S).reshape(-1)), float(a),
                                        int(seed_u) & _U64_MASK, int(off_u) & _U64_MASK, int(seed_n) & _U64_MASK,
                                        int(off_n) & _U64_MASK, 1 if scale_u else 0, _ptr(self._idxs), int(k),
                                        1 if own_mean else 0, _ptr(dst), _ptr(self._invalid), _ptr(self.counter),
                                        _stream()))
        self.launches += 1

    def lw_merge_move(self, mean, S, a, seed_e, off_e, seed_n, off_n, n_new, postselect, dst=None, scale_u=False,
                      u_out=None, js_out=None):
        """Merge draw + move (sorted uniforms from exponential spacings, streaming merge with the CDF)."""
        if dst is None:
            self._alt_slab(n_new)
            dst = self.x_alt
        self._fused_scratch(n_new)
        if self._parent_inv is None or self._parent_inv.numel() < n_new:
            self._parent_inv = torch.empty((n_new,), dtype=torch.int32, device=self.device)
        check(self.lib.qb_lw_merge_move(self.lib_model, _ptr(self.x), self.n, self.d, _ptr(self._cdf), _ptr(self.ws),
                                        self.ws_bytes, 1, _lib.f64_array(mean),
                                        _lib.f64_array(np.asarray(S).reshape(-1)), float(a),
                                        int(seed_e) & _U64_MASK, int(off_e) & _U64_MASK, int(seed_n) & _U64_MASK,
                                        int(off_n) & _U64_MASK, 1 if scale_u else 0, int(n_new), _ptr(dst),
                                        int(bool(postselect)), _ptr(self._invalid), _ptr(self._parent_inv),
                                        _ptr(self.counter), _ptr(u_out) if u_out is not None else None,
                                        _ptr(js_out) if js_out is not None else None, _stream()))
        self.launches += 3

    def lw_merge_retry(self, mean, S, a, seed_n, off_n, k, dst=None):
        dst = self.x_alt if dst is None else dst
        check(self.lib.qb_lw_merge_retry(self.lib_model, _ptr(self.x), self.n, self.d, _lib.f64_array(mean),
                                         _lib.f64_array(np.asarray(S).reshape(-1)), float(a),
                                         int(seed_n) & _U64_MASK, int(off_n) & _U64_MASK, _ptr(self._idxs), int(k),
                                         _ptr(self._parent_inv), _ptr(dst), _ptr(self._invalid), _ptr(self.counter),
                                         _stream()))
        self.launches += 1

    # ---- binned resample (device RNG, d <= 4): csrc/qb_binned.cu ------------------------------------------------
    def binned_supported(self, n_new):
        return self.d <= 4 and self.n < (1 << 31) and 1 <= int(n_new) < (1 << 31)

    def _binned_scratch(self, n_new):
        """Workspace and invalid list for a cloud of up to ``capacity`` particles resampled into up to ``n_new``."""
        if self._bin_ws is None or self._bin_cap[0] < self.capacity or self._bin_cap[1] < n_new:
            cap_new = max(int(n_new), self.capacity, self._bin_cap[1])
            nbytes = self.lib.qb_lw_binned_workspace_bytes(self.capacity, cap_new)
            self._bin_ws = torch.zeros(((nbytes + 7) // 8,), dtype=torch.float64, device=self.device)
            self._bin_cap = (self.capacity, cap_new)
            self._bin_list = torch.empty((cap_new,), dtype=torch.int64, device=self.device)
            self._bin_parents = torch.empty((cap_new,), dtype=torch.int32, device=self.device)
        if self._bin_mirror is None:
            self._bin_mirror = torch.zeros((64,), dtype=torch.float64, pin_memory=True)
            self._bin_mirror_np = self._bin_mirror.numpy()

    def preallocate_binned(self, n_new=None):
        self._binned_scratch(self.n if n_new is None else int(n_new))
        self.preallocate_resample_slab()

    def binned_prepare(self, n_new, seed_u, off_u):
        """Queue pass 1 + 2 of the binned resample: bin sums, moments (device buffer + pinned mirror) and the
        multinomial counts of ``n_new`` draws.  Returns the tag ``binned_moments_wait`` polls for."""
        self._binned_scratch(n_new)
        self._bin_tag += 1
        check(self.lib.qb_lw_binned_prepare(self.x.data_ptr(), self.w.data_ptr(), self.stats.data_ptr(), self.n,
                                            self.d, int(n_new), int(seed_u) & _U64_MASK, int(off_u) & _U64_MASK,
                                            self.count_mode, self.moments_out.data_ptr(), self._bin_mirror.data_ptr(),
                                            float(self._bin_tag), self._bin_ws.data_ptr(), self._bin_ws.numel() * 8,
                                            _stream()))
        self.launches += 2
        return self._bin_tag

    @nvtx_range('qb.cloud.binned_resample')
    def binned_resample(self, n_new, a, h, zero_cov_comp, seed, off_u, off_v, seed_n, off_n, postselect,
                        retry_rounds, fuse_weights, n_global=None, own_mean=False):
        """The whole binned resample queued by one call (moments, Liu-West constants on the device, counts, move,
        first retry launch).  Returns the tag; poll ``binned_moments_wait`` / ``binned_flags`` for the host-side checks
        and ``binned_counters_wait`` / ``binned_retry_wait(queued=True)`` for the result."""
        n_new = int(n_new)
        self._binned_scratch(n_new)
        self._alt_slab(n_new)
        self._bin_tag += 1
        tag = self._bin_tag
        rounds = int(retry_rounds) if (retry_rounds and postselect) else 0
        check(self.lib.qb_lw_binned_resample(
            self.lib_model, self.x.data_ptr(), self.w.data_ptr(), self.stats.data_ptr(), self.n, self.d, n_new,
            float(a), float(h), float(zero_cov_comp), int(seed) & _U64_MASK, int(off_u) & _U64_MASK,
            self.count_mode, int(off_v) & _U64_MASK, int(seed_n) & _U64_MASK, int(off_n) & _U64_MASK, self.x_alt.data_ptr(),
            self.w_alt.data_ptr() if fuse_weights else None, n_new if n_global is None else int(n_global),
            self.stats_alt.data_ptr() if fuse_weights else None, 1 if postselect else 0, rounds,
            1 if own_mean else 0, self._bin_list.data_ptr(), self._bin_parents.data_ptr(),
            self.moments_out.data_ptr(), self._bin_mirror.data_ptr(), float(tag),
            self._bin_ws.data_ptr(), self._bin_ws.numel() * 8, _stream()))
        self.launches += 4 if rounds else 3
        return tag

    # ---- small clouds, parity mode: one single-CTA launch --------------------------------------------------------
    def small_supported(self, n_new):
        return self.d <= 4 and self.n <= _lib.QB_SMALL_MAX and n_new <= 4 * _lib.QB_SMALL_MAX

    def small_resample(self, u, eps, n_new, a, h, zero_cov_comp, postselect, fuse_weights):
        """Queue the whole first Liu-West pass of a small cloud (csrc/qb_binned.cu, lw_small_resample_kernel) on the
        HOST-drawn variates ``u`` (n_new,) and ``eps`` (d, n_new): ONE pinned staging copy, ONE launch.  Returns the
        tag ``binned_moments_wait`` / ``binned_flags`` / ``binned_counters_wait`` poll for; ``small_consts`` then
        gives the constants the device used (for the staged retry kernels)."""
        n_new, d = int(n_new), self.d
        self._resample_scratch(n_new)
        self._alt_slab(n_new)
        if self._bin_mirror is None:
            self._bin_mirror = torch.zeros((64,), dtype=torch.float64, pin_memory=True)
            self._bin_mirror_np = self._bin_mirror.numpy()
        stage = getattr(self, '_small_stage', None)
        if stage is None or stage.numel() < (1 + d) * n_new:
            self._small_stage = stage = torch.empty(((1 + d) * n_new,), dtype=torch.float64, pin_memory=True)
            self._small_stage_np = stage.numpy()
            self._small_dev = torch.empty(((1 + d) * n_new,), dtype=torch.float64, device=self.device)
        sn = self._small_stage_np
        sn[:n_new] = u
        sn[n_new:(1 + d) * n_new] = eps.reshape(-1)
        dev = self._small_dev
        dev[:(1 + d) * n_new].copy_(stage[:(1 + d) * n_new], non_blocking=True)
        self._bin_tag += 1
        tag = self._bin_tag
        base = dev.data_ptr()
        check(self.lib.qb_lw_small_resample(
            self.lib_model, self.x.data_ptr(), self.w.data_ptr(), self.stats.data_ptr(), self.n, d, float(a), float(h),
            float(zero_cov_comp), base, base + 8 * n_new, n_new, self.x_alt.data_ptr(), self._js.data_ptr(),
            self._invalid.data_ptr(), self.counter.data_ptr(), self.w_alt.data_ptr() if fuse_weights else None,
            self.stats_alt.data_ptr() if fuse_weights else None, 1 if postselect else 0, self.moments_out.data_ptr(),
            self._bin_mirror.data_ptr(), float(tag), _stream()))
        self.launches += 1
        return tag

    def small_consts(self):
        """(S (d, d) scaled by h) the last ``small_resample`` derived on the device (valid after its moments tag)."""
        d = self.d
        return self._bin_mirror_np[40:40 + d * d].reshape(d, d).copy()

    def binned_flags(self):
        """(covariance flag, sqrtm error) published with the moments of the last ``binned_resample``."""
        return int(self._bin_mirror_np[29]), float(self._bin_mirror_np[30])

    def binned_sums(self, mirror=True):
        """Pass 1 alone (sharded clouds: the shard masses decide this slab's offspring count)."""
        self._binned_scratch(self.n)
        self._bin_tag += 1
        check(self.lib.qb_lw_binned_sums(_ptr(self.x), _ptr(self.w), _ptr(self.stats), self.n, self.d,
                                         _ptr(self.moments_out),
                                         ctypes.c_void_p(self._bin_mirror.data_ptr()) if mirror else None,
                                         float(self._bin_tag), _ptr(self._bin_ws), self._bin_ws.numel() * 8,
                                         _stream()))
        self.launches += 1
        return self._bin_tag

    def binned_shard_consts(self, rows, n_ranks, a, h, zero_cov_comp):
        """Sharded clouds: global moments (``rows`` = the all-gathered moment sums, device) and the Liu-West constants
        derived on the device; a following ``binned_move(None, None, ...)`` uses them.  Returns the tag
        ``binned_moments_wait`` / ``binned_flags`` poll for."""
        self._bin_tag += 1
        check(self.lib.qb_lw_binned_shard_consts(_ptr(rows), int(n_ranks), self.d, float(a), float(h),
                                                 float(zero_cov_comp), ctypes.c_void_p(self._bin_mirror.data_ptr()),
                                                 float(self._bin_tag), _ptr(self._bin_ws), self._bin_ws.numel() * 8,
                                                 _stream()))
        self.launches += 1
        return self._bin_tag

    def binned_count(self, n_new, seed_u, off_u):
        """Pass 2 alone: multinomial counts of ``n_new`` draws over the bins of the last ``binned_sums``."""
        if self._bin_cap[1] < n_new:
            raise _lib.QbError("binned_count: call binned_reserve(n_new) before binned_sums")
        check(self.lib.qb_lw_binned_count(self.n, int(n_new), int(seed_u) & _U64_MASK, int(off_u) & _U64_MASK,
                                          self.count_mode, _ptr(self._bin_ws), self._bin_ws.numel() * 8, _stream()))
        self.launches += 1

    def _spin(self, index, tag, what, timeout_s=120.0):
        m = self._bin_mirror_np
        want = float(tag)
        spins, t0 = 0, None
        while m[index] != want:
            spins += 1
            if spins > 20000:
                if t0 is None:
                    t0 = time.perf_counter()
                elif time.perf_counter() - t0 > timeout_s:
                    torch.cuda.synchronize()
                    raise _lib.QbError("timed out waiting for %s (tag %d)" % (what, tag))

    def binned_moments_wait(self, tag):
        """(sum w, mean (d,), second moment (d, d)) published by ``binned_prepare``'s first kernel."""
        self._spin(31, tag, "the binned resample's moments")
        d = self.d
        out = self._bin_mirror_np[:1 + d + d * d].copy()
        return out[0], out[1:1 + d], out[1 + d:].reshape(d, d)

    def _bin_consts(self, mean, S):
        """mean (d,) and S (d, d) as the ctypes arrays the C ABI takes (allocated once, refilled)."""
        if getattr(self, '_bin_mean_c', None) is None:
            self._bin_mean_c = (ctypes.c_double * 4)()
            self._bin_S_c = (ctypes.c_double * 16)()
        if mean is None:                 # constants derived on the device (binned_resample)
            return None, None
        d = self.d
        mc, sc = self._bin_mean_c, self._bin_S_c
        if d == 1:
            mc[0] = float(mean[0])
            sc[0] = float(S[0][0]) if not isinstance(S, float) else S
        else:
            flat = np.asarray(S, dtype=np.float64).reshape(-1)
            for c in range(d):
                mc[c] = float(mean[c])
            for j in range(d * d):
                sc[j] = float(flat[j])
        return mc, sc

    def binned_move(self, mean, S, a, seed_v, off_v, seed_n, off_n, n_new, postselect, dst=None, split=None,
                    dst2=None, fuse_weights=False, n_global=None, js_out=None, retry_rounds=0, own_mean=False):
        """Pass 3: every output slot draws inside its bin, moves and is tested; with ``fuse_weights`` the new uniform
        weights and their stats block go to the alternate weight buffer in the same launch.  ``retry_rounds`` > 0
        (postselection only) queues the retry kernel right behind it: up to that many fresh perturbations per invalid
        slot, round j drawing from the normal stream at ``off_n + (j + 1) * stride``, stride = (d n_new + 1) // 2.
        Returns the tag(s) ``binned_counters_wait`` / ``binned_retry_wait`` poll for."""
        if dst is None:
            self._alt_slab(n_new)
            dst = self.x_alt
        self._bin_tag += 1
        tag = self._bin_tag
        n_new = int(n_new)
        n_global = n_new if n_global is None else int(n_global)
        w_new = self.w_alt.data_ptr() if fuse_weights else None
        st_new = self.stats_alt.data_ptr() if fuse_weights else None
        mc, sc = self._bin_consts(mean, S)
        stream = _stream()
        mirror = self._bin_mirror.data_ptr()
        lib, model, d = self.lib, self.lib_model, self.d
        x, split = self.x.data_ptr(), int(n_new if split is None else split)
        dst_p, dst2_p = dst.data_ptr(), (dst2.data_ptr() if dst2 is not None else None)
        rounds = int(retry_rounds) if (retry_rounds and postselect) else 0
        check(lib.qb_lw_binned_move(model, x, self.w.data_ptr(), self.stats.data_ptr(), self.n, d, mc, sc, float(a),
                                    int(seed_v) & _U64_MASK, int(off_v) & _U64_MASK, int(seed_n) & _U64_MASK,
                                    int(off_n) & _U64_MASK, n_new, dst_p, split, dst2_p, w_new, n_global, st_new,
                                    1 if postselect else 0, rounds, 1 if own_mean else 0,
                                    self._bin_list.data_ptr(), self._bin_parents.data_ptr(),
                                    js_out.data_ptr() if js_out is not None else None, mirror + 32 * 8, float(tag),
                                    self._bin_ws.data_ptr(), self._bin_ws.numel() * 8, stream))
        self.launches += 2 if rounds else 1
        return (tag, tag) if rounds else tag

    def binned_retry(self, mean, S, a, seed_n, off_n, n_new, rounds=1, dst=None, split=None, dst2=None,
                     own_mean=False, seed_v=0):
        """Up to ``rounds`` more perturbations per still-invalid slot of the last move (round j: normal stream at
        ``off_n + j * stride``)."""
        dst = self.x_alt if dst is None else dst
        self._bin_tag += 1
        mc, sc = self._bin_consts(mean, S)
        n_new = int(n_new)
        stride = (self.d * n_new + 1) // 2
        check(self.lib.qb_lw_binned_retry(self.lib_model, self.x.data_ptr(), self.n, self.d, mc, sc, float(a),
                                          int(seed_n) & _U64_MASK, int(off_n) & _U64_MASK, stride, int(rounds),
                                          1 if own_mean else 0, int(seed_v) & _U64_MASK, n_new,
                                          dst.data_ptr(), int(n_new if split is None else split),
                                          dst2.data_ptr() if dst2 is not None else None, self._bin_list.data_ptr(),
                                          self._bin_parents.data_ptr(), self._bin_mirror.data_ptr() + 40 * 8,
                                          float(self._bin_tag), self._bin_ws.data_ptr(), self._bin_ws.numel() * 8,
                                          _stream()))
        self.launches += 1
        return self._bin_tag

    def binned_counters_wait(self, tag):
        """(#invalid, #clamped draws, #drawn) of the move launch ``tag`` (from the pinned mirror)."""
        self._spin(32 + 3, tag, "the binned resample's move kernel")
        m = self._bin_mirror_np
        return int(m[32]), int(m[33]), int(m[34])

    @nvtx_range('qb.cloud.binned_retry_wait')
    def binned_retry_wait(self, tag, queued=False):
        """(#still invalid, most rounds used, list length) of a retry launch: ``queued`` = the one ``binned_move`` queued
        behind itself (it reports next to the move's block, under the move's tag)."""
        base = 36 if queued else 40
        self._spin(base + 3, tag, "the binned resample's retry kernel")
        m = self._bin_mirror_np
        return int(m[base]), int(m[base + 1]), int(m[base + 2])

    def adopt_binned(self, n_new, weights_fused, n_global=None):
        """Make the slab the binned move wrote current.  With fused weights the alternate weight/stats buffers already
        hold 1/n and its stats block: flip the ping-pong index instead of launching the fill kernel."""
        if not weights_fused:
            return self.adopt_resampled(n_new, n_global)
        self._swap_slabs(n_new)
        self.cur = 1 - self.cur
        self._chain_tag = 0
        self._slot_tag = [0, 0]

    def read_counter(self):
        self.counter_host.copy_(self.counter[:2], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(self.counter_host[0]), int(self.counter_host[1])

    def compact_invalid(self, n_new):
        check(self.lib.qb_compact_invalid(_ptr(self._invalid), int(n_new), _ptr(self._idxs), _ptr(self.counter),
                                          _ptr(self.ws), self.ws_bytes, _stream()))
        self.launches += 3

    def lw_retry(self, mean, S, a, eps_dev, k, x_src=None, own_mean=False, js=None):
        src = self.x if x_src is None else x_src
        check(self.lib.qb_lw_retry(self.lib_model, _ptr(src), src.shape[0], self.d,
                                   _ptr(self._js if js is None else js),
                                   _ptr(self._idxs), int(k), _lib.f64_array(mean),
                                   _lib.f64_array(np.asarray(S).reshape(-1)), float(a), _ptr(eps_dev),
                                   _ptr(self.x_alt), _ptr(self._invalid), _ptr(self.counter),
                                   1 if own_mean else 0, _ptr(self._lw_consts()), self._lw_consts().numel() * 8,
                                   _stream()))
        self.launches += 1

    def _swap_slabs(self, n_new):
        """The alternate slab (holding ``n_new`` new particles) becomes current; every per-particle buffer follows the
        new particle count.  Within ``capacity`` nothing is allocated; beyond it the weight buffers and the
        workspaces are rebuilt for the new size (a resampler asked for more particles, resamplers.py:86-88)."""
        n_new = int(n_new)
        self._x_back, self._x_alt_back = self._x_alt_back, self._x_back
        if n_new > self.capacity or n_new > self._w_back[0].numel():
            self.capacity = n_new
            f64 = dict(dtype=torch.float64, device=self.device)
            self._w_back = [torch.empty((n_new,), **f64), torch.empty((n_new,), **f64)]
            self._alloc_workspace()
            self._bin_ws = None
            self._cdf = self._js = self._u = self._eps = self._invalid = self._idxs = self._parent_inv = None
            if self._x_alt_back is not None and self._x_alt_back.shape[0] < n_new:
                self._x_alt_back = None
        self.n = n_new
        self.x = self._x_back[:n_new]
        self.x_alt = self._x_alt_back[:n_new] if self._x_alt_back is not None else None
        self._w = [b[:n_new] for b in self._w_back]
        self._chain_tag = 0

    def resize(self, n_new):
        """Change the particle count within ``capacity`` without touching the data (the caller refills the slab)."""
        n_new = int(n_new)
        if n_new > self.capacity:
            raise _lib.QbError("resize(%d) beyond the cloud's capacity %d" % (n_new, self.capacity))
        self.n = n_new
        self.x = self._x_back[:n_new]
        self.x_alt = None
        self._w = [b[:n_new] for b in self._w_back]
        self._chain_tag = 0
        self._slot_tag = [0, 0]

    def adopt_resampled(self, n_new, n_global=None):
        """Make the freshly written slab current; weights become uniform."""
        self._swap_slabs(n_new)
        self.set_uniform_weights(n_global)

    @nvtx_range('qb.cloud.canonicalize')
    def canonicalize(self):
        if self.desc.kind != _lib.QB_MODEL_TOMOGRAPHY:
            return
        if self.n >= 4096:
            # large clouds: LDL^H screening pass, eigendecomposition only for the particles it cannot certify
            self._fused_scratch(self.n)
            if getattr(self, '_canon_count', None) is None:
                self._canon_count = torch.zeros((1,), dtype=torch.int64, device=self.device)
            check(self.lib.qb_tomo_canonicalize_screened_ld(_ptr(self.x), self.n, self.desc.dim, self.d,
                                                            _ptr(self.basis_dev), int(self.desc.allow_subnormalized),
                                                            _ptr(self._invalid), _ptr(self._idxs),
                                                            _ptr(self._canon_count), _ptr(self.ws), self.ws_bytes,
                                                            _stream()))
            self.launches += 5
            return
        check(self.lib.qb_tomo_canonicalize_ld(_ptr(self.x), self.n, self.desc.dim, self.d, _ptr(self.basis_dev),
                                               int(self.desc.allow_subnormalized), _stream()))
        self.launches += 1

    # ---- read-side estimators on the device -------------------------------------------------------
    def _readside_ws(self):
        ws = getattr(self, '_rs_ws', None)
        if ws is None:
            nbytes = self.lib.qb_readside_workspace_bytes()
            self._rs_ws = ws = torch.zeros(((nbytes + 7) // 8,), dtype=torch.float64, device=self.device)
            self._rs_mass = torch.zeros((2048,), dtype=torch.float64, device=self.device)
            self._rs_count = torch.zeros((2048,), dtype=torch.int64, device=self.device)
            self._rs_out = torch.zeros((1,), dtype=torch.float64, device=self.device)
        return ws

    def entropy(self):
        """-sum_{w > 0} w log w of the normalised weights (distributions.py:457-465); 8 bytes come back."""
        ws = self._readside_ws()
        check(self.lib.qb_weights_entropy(_ptr(self.w), _ptr(self.stats), self.n, _ptr(self._rs_out), _ptr(ws),
                                          ws.numel() * 8, _stream()))
        self.launches += 2
        return float(self._rs_out.cpu().numpy()[0])

    def credible_members(self, level):
        """Indices (device int64 tensor) of the K heaviest particles, K = 1 + #{k : cumsum of the descending weights
        <= level} (distributions.py:583-592), found by radix selection on the weights' bit patterns — no sort, and
        only 6 x 32 KB of histograms travel to the host.  Equal weights at the threshold are taken lowest index
        first (the reference's order among ties is that of an unstable argsort)."""
        self._readside_ws()
        n = self.n
        prefix, mass_above, count_above = 0, 0.0, 0
        last = None
        for shift, nb in ((53, 11), (42, 11), (31, 11), (20, 11), (9, 11), (0, 9)):
            check(self.lib.qb_weight_mass_hist(_ptr(self.w), _ptr(self.stats), n, shift, nb, prefix,
                                               _ptr(self._rs_mass), _ptr(self._rs_count), _stream()))
            self.launches += 1
            nbk = 1 << nb
            mass = self._rs_mass.cpu().numpy()[:nbk][::-1]           # from the heaviest bucket down
            count = self._rs_count.cpu().numpy()[:nbk][::-1]
            incl = mass_above + np.cumsum(mass)
            over = np.nonzero((incl > level) & (count > 0))[0]
            if over.size == 0:
                # every cumulative weight is <= level: the reference indexes one past the end (distributions.py:591)
                raise IndexError("index %d is out of bounds for axis 0 with size %d" % (n, n))
            j = int(over[0])
            mass_above += float(np.sum(mass[:j]))
            count_above += int(np.sum(count[:j]))
            prefix = (prefix << nb) | (nbk - 1 - j)
            last = (float(mass[j]), int(count[j]))
        tau = prefix
        v = float(np.array([tau], dtype=np.uint64).view(np.float64)[0])
        c_eq = last[1]
        # how many of the c_eq particles weighing exactly tau still fit under `level`: the reference's np.cumsum adds
        # them one by one, so do the same (from the mass above them) instead of multiplying
        k_est = int(np.floor((level - mass_above) / v)) if v > 0 else 0
        m = int(min(c_eq, max(k_est, 0) + 8))
        run = np.cumsum(np.concatenate([[mass_above], np.full(m, v)]))[1:]
        k_in = int(np.sum(run <= level))
        take = min(k_in + 1, c_eq)
        # members: everything heavier than tau, plus `take` of the particles that weigh exactly tau
        flags = torch.empty((n,), dtype=torch.uint8, device=self.device)
        cnt = torch.zeros((1,), dtype=torch.int64, device=self.device)
        parts = []
        for mode, want in ((0, count_above), (1, take)):
            if want == 0:
                continue
            idx = torch.empty((n,), dtype=torch.int64, device=self.device)
            check(self.lib.qb_weights_select(_ptr(self.w), _ptr(self.stats), n, tau, mode, _ptr(flags), _stream()))
            check(self.lib.qb_compact_invalid(_ptr(flags), n, _ptr(idx), _ptr(cnt), _ptr(self.ws), self.ws_bytes,
                                              _stream()))
            self.launches += 4
            parts.append(idx[:want])
        return torch.cat(parts) if len(parts) > 1 else parts[0]

    def gather_members(self, idx):
        """(locations (k, d), normalised weights (k,)) of the particles ``idx`` as host arrays: 8 k (d + 1) bytes."""
        k = idx.numel()
        rows = torch.empty((k, self.d), dtype=torch.float64, device=self.device)
        wts = torch.empty((k,), dtype=torch.float64, device=self.device)
        check(self.lib.qb_gather_rows(_ptr(self.x), self.d, _ptr(idx), k, _ptr(rows), _stream()))
        check(self.lib.qb_gather_rows(_ptr(self.w), 1, _ptr(idx), k, _ptr(wts), _stream()))
        self.launches += 2
        inv = float(self.read_stats()[QB_STAT_INV_NORM])
        self.last_d2h_bytes = 8 * k * (self.d + 1)
        return self._to_host(rows), self._to_host(wts) * inv

    def weighted_mean_of(self, values):
        """sum_i w_i values[i, :] for an (n, k) float64 device tensor (est_meanfn, distributions.py:411-430), through
        the moment kernel, k <= 64 columns per launch; 8 k bytes come back."""
        n, k = values.shape
        out = np.empty((k,))
        for c0 in range(0, k, _lib.QB_MAX_D):
            c1 = min(k, c0 + _lib.QB_MAX_D)
            blk = values[:, c0:c1].contiguous()
            kk = c1 - c0
            need = self.lib.qb_moments_workspace_bytes(n, kk)
            ws = getattr(self, '_meanfn_ws', None)
            if ws is None or ws.numel() * 8 < need:
                self._meanfn_ws = ws = torch.zeros(((need + 7) // 8,), dtype=torch.float64, device=self.device)
            res = torch.empty((1 + kk + kk * kk,), dtype=torch.float64, device=self.device)
            check(self.lib.qb_moments(_ptr(blk), _ptr(self.w), _ptr(self.stats), n, kk, _ptr(res), _ptr(ws),
                                      ws.numel() * 8, _stream()))
            self.launches += 2
            out[c0:c1] = res[1:1 + kk].cpu().numpy()
        return out

    # ---- f4 decorators: Gaussian steps after an update, Gaussian noise on the likelihood -------------------------
    def normals(self, count):
        """``count`` standard normals in a device buffer, from the cloud's noise source (set by the updater):
        'numpy' draws np.random.normal on the host exactly where the reference does and uploads, 'mt19937' continues
        the same legacy stream on the device, 'philox' is the counter-based device generator."""
        count = int(count)
        buf = getattr(self, '_noise_buf', None)
        if buf is None or buf.numel() < count:
            self._noise_buf = buf = torch.empty((count,), dtype=torch.float64, device=self.device)
        buf = buf[:count]
        kind = self.noise_rng
        if kind == 'numpy':
            buf.copy_(torch.from_numpy(np.random.normal(size=(count,))))
        elif kind == 'mt19937':
            self.mt19937_normal(buf, count)
        else:
            self.rng_normal(buf, count, self.noise_seed, self.noise_offset)
            self.noise_offset += (count + 1) // 2
        return buf

    def walk_step(self, expparams):
        """Model.update_timestep of the decorator described by ``desc.walk`` (derived_models.py:733-741, 921-963;
        tomography/models.py:257-272), in place on the committed slab."""
        walk = self.desc.walk
        n, d, kind = self.n, self.d, walk['kind']
        I32 = ctypes.c_int32
        mult = 1.0
        if kind == 'generic':                               # the model's own step distribution: a host object
            steps = np.ascontiguousarray(walk['dist'].sample(n=n), dtype=np.float64).reshape(n, -1)
            k = steps.shape[1]
            z = torch.from_numpy(steps).to(self.device).reshape(-1)
            idx = zcol = list(range(k))
            mode, scale, sidx, pre = _lib.QB_WALK_ADD, None, None, 1.0
        elif kind == 'diffusive':                           # eps * sqrt(t) * randn, first and last parameter fixed
            k = d
            z = self.normals(n * d)
            idx = zcol = list(range(1, d - 1))
            mode, scale, sidx = _lib.QB_WALK_LEARNED, None, [d - 1] * (d - 2)
            pre = float(np.sqrt(np.asarray(expparams)['t'].reshape(-1)[0]))
        else:
            idx = walk['idxs']
            k = len(idx)
            zcol = list(range(k))
            fn = walk.get('scale_mult')
            if fn is not None:
                mult = float(np.ravel(np.asarray(fn(expparams), dtype=float))[0])
            pre = 1.0
            if kind == 'fixed':
                z, mode, scale, sidx = self.normals(n * k), _lib.QB_WALK_FIXED, walk['scale'], None
            elif kind == 'learned':
                z, mode, scale, sidx = self.normals(n * k), _lib.QB_WALK_LEARNED, None, walk['sidx']
            else:                                           # fixed dense covariance: chol @ normal(n_rw, n) on the host
                steps = np.dot(walk['chol'], np.random.normal(size=(k, n))).T
                z = torch.from_numpy(np.ascontiguousarray(steps)).to(self.device).reshape(-1)
                mode, scale, sidx = _lib.QB_WALK_ADD, None, None
        n_rw = len(idx)
        check(self.lib.qb_walk_step(_ptr(self.x), n, d, n_rw, (I32 * n_rw)(*idx), (I32 * n_rw)(*zcol), mode,
                                    _lib.f64_array(scale) if scale is not None else None,
                                    (I32 * n_rw)(*sidx) if sidx is not None else None, pre, mult, _ptr(z), k,
                                    _stream()))
        self.launches += 1
        if kind == 'diffusive':
            self.canonicalize()

    def _poisoned_update(self, ep, outcome, src, dst, ctl, stream):
        """PoisonedModel (derived_models.py:188-204): plain likelihood -> + clipped Gaussian noise -> the fused update
        kernel on the poisoned likelihood vector (a one-parameter 'coin' whose pr0 is the stored value)."""
        n, po = self.n, self.desc.poison
        L = getattr(self, '_like_buf', None)
        if L is None or L.numel() < n:
            self._like_buf = L = torch.empty((self.capacity,), dtype=torch.float64, device=self.device)
            self._coin = _lib.QbModel(kind=_lib.QB_MODEL_COIN, d=1, binomial=0, interleaved=0, min_freq=0.0,
                                      likelihood_power=1.0, d_extra=0, extra_rule=0)
            self._coin_ep = _lib.QbExpparams()
        outs = (ctypes.c_int64 * 1)(int(outcome))
        check(self.lib.qb_likelihood(self.lib_model, ctypes.byref(ep), 1, outs, 1, self._px, n, _ptr(L), stream))
        z = self.normals(n)
        check(self.lib.qb_poison_likelihood(_ptr(L), n, _ptr(z), po['mode'], po['tol'], po['denom'], stream))
        check(self.lib.qb_fused_update(ctypes.byref(self._coin), ctypes.byref(self._coin_ep), 0, _ptr(L), n,
                                       self._pw[src], self._pw[dst], self._pstats[src], self._pstats[dst],
                                       ctypes.byref(ctl), self._pws, self.ws_bytes, stream))
        self.launches += 2

    def rng_uniform(self, out, n, seed, offset):
        check(self.lib.qb_rng_uniform(_ptr(out), int(n), int(seed), int(offset), _stream()))
        self.launches += 1

    def rng_normal(self, out, n, seed, offset):
        check(self.lib.qb_rng_normal(_ptr(out), int(n), int(seed), int(offset), _stream()))
        self.launches += 1

    # NumPy's legacy global MT19937 stream continued on the device (parity mode at scale): the state is read
    # from / written back to np.random, so host draws before and after interleave exactly as in the reference.
    def _mt_ws(self, nbytes):
        ws = getattr(self, '_mt_workspace', None)
        if ws is None or ws.numel() < nbytes:
            self._mt_workspace = ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.x.device)
        return ws

    def mt19937_uniform(self, out, n):
        kind, key, pos, has_gauss, cached = np.random.get_state()
        if kind != 'MT19937':
            raise _lib.QbError("np.random is not the legacy MT19937 generator")
        key = np.ascontiguousarray(key, dtype=np.uint32)
        key_out = np.empty(624, dtype=np.uint32)
        pos_out = ctypes.c_int32(0)
        nbytes = self.lib.qb_mt19937_workspace_bytes(int(n), 0)
        ws = self._mt_ws(nbytes)
        check(self.lib.qb_mt19937_uniform(key.ctypes.data, int(pos), int(n), _ptr(out), key_out.ctypes.data,
                                          ctypes.byref(pos_out), _ptr(ws), ws.numel(), _stream()))
        np.random.set_state((kind, key_out, int(pos_out.value), has_gauss, cached))
        self.launches += 2

    def mt19937_normal(self, out, m):
        kind, key, pos, has_gauss, cached = np.random.get_state()
        if kind != 'MT19937':
            raise _lib.QbError("np.random is not the legacy MT19937 generator")
        key = np.ascontiguousarray(key, dtype=np.uint32)
        key_out = np.empty(624, dtype=np.uint32)
        pos_out, hg_out, cached_out = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_double(0.0)
        nbytes = self.lib.qb_mt19937_workspace_bytes(0, int(m))
        ws = self._mt_ws(nbytes)
        check(self.lib.qb_mt19937_normal(key.ctypes.data, int(pos), int(has_gauss), float(cached), int(m), _ptr(out),
                                         key_out.ctypes.data, ctypes.byref(pos_out), ctypes.byref(hg_out),
                                         ctypes.byref(cached_out), _ptr(ws), ws.numel(), _stream()))
        np.random.set_state((kind, key_out, int(pos_out.value), int(hg_out.value), float(cached_out.value)))
        self.launches += 6


# ---------------------------------------------------------------------------
# Host-array conveniences used by the Model classes (upload, run kernel, download)
# ---------------------------------------------------------------------------

def device_likelihood(desc, x_dev, outcomes, expparams):
    """L[o, i, e] for particles already in HBM (``x_dev``: (n, d) float64 cuda tensor)."""
    lib = _lib.load()
    outcomes = np.atleast_1d(np.asarray(outcomes)).astype(np.int64)
    expparams = np.atleast_1d(np.asarray(expparams))
    n, n_e, n_o = x_dev.shape[0], expparams.shape[0], outcomes.shape[0]
    eps = (_lib.QbExpparams * n_e)(*[desc.expparams_record(expparams, e) for e in range(n_e)])
    outs = (ctypes.c_int64 * n_o)(*[int(o) for o in outcomes])
    L = torch.empty((n_o, n, n_e), dtype=torch.float64, device=x_dev.device)
    check(lib.qb_likelihood(ctypes.pointer(desc.c_model), eps, n_e, outs, n_o, _ptr(x_dev), n, _ptr(L), _stream()))
    return L.cpu().numpy()


def host_likelihood(desc, outcomes, modelparams, expparams):
    dev = _require_cuda()
    x = np.ascontiguousarray(modelparams, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    if x.shape[1] != desc.d:
        raise ValueError("modelparams has %d columns, model has %d parameters" % (x.shape[1], desc.d))
    return device_likelihood(desc, torch.from_numpy(x).to(dev), outcomes, expparams)


def host_are_models_valid(desc, modelparams):
    lib = _lib.load()
    dev = _require_cuda()
    x = np.ascontiguousarray(modelparams, dtype=np.float64)
    xd = torch.from_numpy(x).to(dev)
    out = torch.empty((x.shape[0],), dtype=torch.uint8, device=dev)
    check(lib.qb_are_models_valid(ctypes.pointer(desc.c_model), _ptr(xd), x.shape[0], _ptr(out), _stream()))
    return out.cpu().numpy().astype(bool)


def host_canonicalize(desc, modelparams):
    lib = _lib.load()
    dev = _require_cuda()
    x = np.ascontiguousarray(modelparams, dtype=np.float64)
    xd = torch.from_numpy(x).to(dev)
    b = np.ascontiguousarray(desc.basis).view(np.float64).reshape(-1)
    bd = torch.from_numpy(b.copy()).to(dev)
    check(lib.qb_tomo_canonicalize_ld(_ptr(xd), x.shape[0], desc.dim, x.shape[1], _ptr(bd),
                                      int(desc.allow_subnormalized), _stream()))
    return xd.cpu().numpy()


def host_weight_stats(weights):
    """(sum w, sum w^2) of host weights, reduced on the device."""
    lib = _lib.load()
    dev = _require_cuda()
    w = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
    n = w.numel()
    stats = torch.zeros((QB_STAT_COUNT,), dtype=torch.float64, device=dev)
    ws_bytes = lib.qb_update_workspace_bytes(n, 1)
    ws = torch.zeros(((ws_bytes + 7) // 8,), dtype=torch.float64, device=dev)
    check(lib.qb_weights_restat(_ptr(w), n, _ptr(stats), _ptr(ws), ws.numel() * 8, _stream()))
    s = stats.cpu().numpy()
    return float(s[QB_STAT_NORM]), float(s[QB_STAT_SUMSQ])


def host_moments(weights, locations):
    """(sum w, mean, second moment) of a host-held particle set, reduced on the device."""
    lib = _lib.load()
    dev = _require_cuda()
    x = np.ascontiguousarray(locations, dtype=np.float64)
    n, d = x.shape
    xd = torch.from_numpy(x).to(dev)
    w = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)
    stats = torch.zeros((QB_STAT_COUNT,), dtype=torch.float64, device=dev)
    stats[QB_STAT_INV_NORM] = 1.0
    ws_bytes = lib.qb_moments_workspace_bytes(n, d)
    ws = torch.zeros(((ws_bytes + 7) // 8,), dtype=torch.float64, device=dev)
    out = torch.empty((1 + d + d * d,), dtype=torch.float64, device=dev)
    check(lib.qb_moments(_ptr(xd), _ptr(w), _ptr(stats), n, d, _ptr(out), _ptr(ws), ws.numel() * 8, _stream()))
    o = out.cpu().numpy()
    return o[0], o[1:1 + d].copy(), o[1 + d:].reshape(d, d).copy()
