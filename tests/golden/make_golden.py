"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz`` holding (a) the seeded inputs and (b) the outputs
the reference (QInfer @ 8170c84, NumPy 2.3.5 / SciPy 1.18.1, Python 3.12.3)
produced for them through its public API (SMCUpdater.update / batch_update,
LiuWestResampler, Model.likelihood, TomographyModel.canonicalize,
ParticleDistribution moments, utils.sqrtm_psd).  It also replays each recorded
resample event with the NumPy calls the reference makes (resamplers.py:308-321)
to store the reference's CDF and resample indices, and finally checks that the
oracle restatement reproduces every reference output bit for bit.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

import cases  # noqa: E402


def add_draws(out):
    """For each recorded resample event add the reference's cdf / u / js."""
    for i in range(int(out['n_events'])):
        w = out['ev%d_w' % i]
        np.random.set_state(cases.unpack_rng_state(out, 'ev%d_rng_' % i))
        cdf = np.cumsum(w)
        u = np.random.random((w.shape[0],))
        out['ev%d_u' % i] = u
        out['ev%d_js' % i] = cdf.searchsorted(u, side='right').astype(np.int64)
    return out


def build(ns):
    files = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inp = cases.precession_inputs()
        files['precession_c1'] = dict(inp, **add_draws(cases.run_precession(ns, inp)))
        inp = cases.precession_inputs(n_particles=600, n_updates=40, true_omega=0.45, seed=77)
        files['precession_minfreq'] = dict(inp, **add_draws(cases.run_precession(ns, inp, min_freq=0.3, a=0.9)))
        inp = cases.rb_inputs()
        files['rb_binomial_c3'] = dict(inp, **add_draws(cases.run_rb(ns, inp)))
        basis = np.asarray(ns.pauli_basis(2).data)
        inp = cases.tomography_inputs(basis)
        files['tomography_c4'] = dict(inp, **add_draws(cases.run_tomography(ns, inp)))
        files['likelihood_vectors'] = cases.likelihood_vectors(ns)
        files['canonicalize_vectors'] = cases.canonicalize_vectors(ns)
        files['moment_vectors'] = cases.moment_vectors(ns)
        files['design_vectors'] = cases.design_vectors(ns)
        files['mle_vectors'] = cases.mle_vectors(ns)
        files['random_walk_vectors'] = cases.random_walk_vectors(ns)
        files['diffusive_vectors'] = cases.diffusive_vectors(ns)
        files['mle_design_vectors'] = cases.mle_design_vectors(ns)
    return files


def main():
    ref = build(cases.reference_namespace())
    orc = build(cases.oracle_namespace())
    bad = 0
    for name, d in ref.items():
        for k, v in d.items():
            a, b = np.asarray(v), np.asarray(orc[name][k])
            same = a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
            if not same:
                bad += 1
                err = np.max(np.abs(a - b)) if a.shape == b.shape else 'shape'
                print("MISMATCH %s[%s]: %s" % (name, k, err))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print("wrote %s.npz (%d arrays, %d resample events)" % (name, len(d), int(d.get('n_events', 0))))
    print("oracle vs reference: %s" % ("BIT-EXACT on all arrays" if bad == 0 else "%d mismatches" % bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
