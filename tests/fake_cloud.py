"""A NumPy stand-in for ``qinfer_b200.engine.DeviceCloud`` — TEST DOUBLE, CPU test suite only.

It implements the contract the host control flow of ``qinfer_b200.smc.SMCUpdater`` relies on (ping-pong weight /
stats buffers, per-step stats blocks with the launch tag, the device-side attention / guard / skip protocol of
include/qinfer_b200.h: ``qb_update_ctl``), with the arithmetic of csrc/qb_update.cu's ``publish()`` restated in
NumPy, so that speculation, fusion and roll-back can be exercised without a GPU.  The product never imports it.
"""
import numpy as np

from qinfer_b200 import _lib

EPS = np.spacing(1)
(NORM, SUMSQ, MIN, NBAD, INV_NORM, NESS, TAG, SKIPPED, ATTN) = range(9)


class FakeCloud(object):
    def __init__(self, desc, n, device=None):
        assert desc.kind == _lib.QB_MODEL_PRECESSION and not desc.binomial, "the double evaluates precession only"
        self.desc, self.n, self.d = desc, int(n), 1
        self.device = 'cpu'
        self.x = np.zeros((self.n, 1))
        self._w = [np.zeros(self.n), np.zeros(self.n)]
        self._stats = [np.zeros(16), np.zeros(16)]
        self.cur = 0
        self._blocks = {}                      # slot -> (tag, blocks)
        self._tag = 0
        self.launches = 0
        self.update_launches = 0
        self.resample_events = None
        self._chain_tag = 0
        self.log = []                          # (kind, nsteps, guard, cancelled)

    w = property(lambda self: self._w[self.cur])
    stats = property(lambda self: self._stats[self.cur])

    # ---- host <-> "device" -----------------------------------------------------------------------
    def upload_locations(self, locs):
        self.x = np.array(locs, dtype=float).reshape(self.n, 1)

    def download_locations(self):
        return self.x.copy()

    def upload_weights(self, w):
        self._w[self.cur] = np.array(w, dtype=float)
        st = self._stats[self.cur]
        st[:] = 0
        st[NORM], st[SUMSQ] = np.sum(w), np.sum(np.asarray(w) ** 2)
        st[INV_NORM], st[NESS] = 1.0, 1.0 / st[SUMSQ]

    def _normalised(self, slot):
        # The device multiplies by the stored reciprocal 1/S; the double DIVIDES, like smc.py:373, so that an unfused
        # run is bit-identical to the oracle (the reciprocal's last-bit difference is amplified by t ~ 1e5 late in a run)
        st = self._stats[slot]
        return self._w[slot] / st[NORM] if st[INV_NORM] != 1.0 else self._w[slot].copy()

    def download_weights(self):
        return self._normalised(self.cur)

    def set_uniform_weights(self, n_global=None):
        n_global = self.n if n_global is None else n_global
        v = 1.0 / n_global
        self._w[self.cur] = np.full(self.n, v)
        st = self._stats[self.cur]
        st[:] = 0
        st[NORM], st[SUMSQ], st[MIN], st[INV_NORM], st[NESS] = 1.0, v, v, 1.0, float(n_global)

    def read_stats(self, which=None):
        return (self.stats if which is None else which).copy()

    def canonicalize(self):
        pass

    def moments(self):
        w = self.download_weights()
        mean = np.dot(w, self.x)
        m2 = np.einsum('i,mi,ni', w, self.x.T, self.x.T)
        return np.sum(w), mean, m2

    # ---- the fused update (csrc/qb_update.cu: kernel body + publish()) ---------------------------------
    def fused_update(self, steps, src, guard=False, zero_weight_thresh=0.0, resample_below=0.0):
        dst = 1 - src
        self._tag += 1
        tag = float(self._tag)
        k = len(steps)
        st_in = self._stats[src]
        self.launches += 1
        self.update_launches += 1
        if guard and (st_in[ATTN] != 0.0 or st_in[SKIPPED] != 0.0):
            self._stats[dst][SKIPPED] = 1.0
            self._blocks[dst] = (tag, np.tile(np.array([0, 0, 0, tag, 0, 0, tag, 2.0]), (k, 1)))
            self.log.append(('update', k, guard, True))
            return self._tag
        w = self._normalised(src)
        blocks = np.zeros((k, 8))
        attn, s_prev, nbad_tot = 0.0, 1.0, 0.0
        for j, (ep, outcome, chk) in enumerate(steps):
            pr0 = np.cos(ep.t * (self.x[:, 0] - ep.w_) / 2) ** 2              # test_models.py:134-140
            w = w * (pr0 if outcome == 0 else 1 - pr0)
            S, Q = np.sum(w), np.sum(w * w)
            nb = float(np.sum(~(w >= 0)))
            rec = S if j == 0 else S / s_prev
            degenerate = abs(rec) < EPS
            total = rec if degenerate else 1.0
            with np.errstate(divide='ignore', invalid='ignore'):
                ne = 1.0 / Q if degenerate else (S * S) / Q
            a = (nb > 0) or (total <= zero_weight_thresh)
            if chk:
                a = a or (ne < resample_below)
            if a and attn == 0.0:
                attn = float(j + 1)
            blocks[j] = [S, Q, nb, tag, rec, ne, tag, 1.0 if a else 0.0]
            s_prev, nbad_tot = S, nbad_tot + nb
        self._w[dst] = w
        so = self._stats[dst]
        so[:] = 0
        so[NORM], so[SUMSQ], so[MIN], so[NBAD] = S, Q, np.nan, nbad_tot
        so[INV_NORM] = 1.0 if abs(S) < EPS else 1.0 / S
        so[NESS], so[TAG], so[SKIPPED], so[ATTN] = ne, tag, 0.0, attn
        self._blocks[dst] = (tag, blocks)
        self.log.append(('update', k, guard, False))
        return self._tag

    def wait_stats(self, slot, tag, nsteps=1, timeout_s=120.0):
        t, blocks = self._blocks[slot]
        assert t == float(tag) and blocks.shape[0] == nsteps, "host waited for a launch that is not the pending one"
        return blocks.copy()

    def commit_update(self):
        self.cur = 1 - self.cur

    def pending_min_weight(self, slot):
        return float(np.min(self._w[slot]))

    def clip_weights(self, slot):
        st = self._stats[slot]
        w = np.clip(self._normalised(slot), 0, 1)
        self._w[slot] = w
        st[NORM], st[SUMSQ], st[MIN], st[NBAD], st[INV_NORM] = np.sum(w), np.sum(w * w), np.min(w), 0.0, 1.0
        st[NESS] = 1.0 / st[SUMSQ]
        return st.copy()

    # ---- experiment design (csrc/qb_design.cu restated) ------------------------------------------------
    def design_sums(self, expparams, idx, outcomes, centre, want_kld):
        ep = self.desc.expparams_record(expparams, idx)
        wn = self.download_weights()
        dx = self.x - np.asarray(centre)[None, :]
        n_o = len(outcomes)
        sums, kld = np.zeros((n_o, 1 + 2 * self.d)), np.zeros(n_o)
        pr0 = np.cos(ep.t * (self.x[:, 0] - ep.w_) / 2) ** 2
        for o, outcome in enumerate(outcomes):
            h = wn * (pr0 if outcome == 0 else 1 - pr0)
            sums[o, 0] = np.sum(h)
            sums[o, 1:1 + self.d] = h @ dx
            sums[o, 1 + self.d:] = h @ dx ** 2
            if want_kld:
                div = 1.0 if (o < n_o - 1 and abs(sums[o, 0]) < EPS) else sums[o, 0]
                wh = h / div
                pos = wh > 0
                kld[o] = np.sum(wh[pos] * np.log(wh[pos] / wn[pos]))
        self.launches += 2
        return sums, (kld if want_kld else None)
