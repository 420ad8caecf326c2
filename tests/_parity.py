"""Replay-parity assertion shared by the GPU tests.

north_star: with both sides consuming the identical replayed uniform stream, per-neutron outcomes must agree to 1e-10.
The measured state (tests/parity_sweep.py over 5e7 neutrons per config, profiles/) is ZERO mismatches, so the tests
assert what was measured: no mismatch at all -- or, where draw counts are available, only "branch flips" (a
comparison that rounds differently on the two sides and sends the neutron down another branch: its draw count then
differs) and at most one per million neutrons.  A regression that breaks one neutron in 2000 fails."""
import numpy as np

SAMPLE_TOL = 1e-10
FLIPS_PER_NEUTRON = 1e-6


def match(outs, refs, tol=SAMPLE_TOL):
    """outs/refs: sequences of arrays; the first pair is an energy (relative tolerance), the rest are cosines /
    direction components (absolute)."""
    ok = np.abs(outs[0] - refs[0]) <= tol * np.maximum(np.abs(refs[0]), 1e-300)
    for a, b in zip(outs[1:], refs[1:]):
        ok &= np.abs(a - b) <= tol
    return ok


def assert_replay(outs, refs, nd=None, nd_ref=None, what=""):
    ok = match(outs, refs)
    n = ok.size
    nbad = int((~ok).sum())
    if nd is not None and nd_ref is not None:
        nd = np.asarray(nd).astype(np.uint32)
        flips = nd != np.asarray(nd_ref).astype(np.uint32)
        numeric = int((~ok & ~flips).sum())
        assert numeric == 0, "%s: %d numeric mismatches beyond 1e-10 without a branch flip" % (what, numeric)
        assert int(flips.sum()) <= max(1, int(FLIPS_PER_NEUTRON * n)) and nbad <= max(1, int(FLIPS_PER_NEUTRON * n)), \
            "%s: %d mismatches / %d draw-count flips in %d neutrons" % (what, nbad, int(flips.sum()), n)
        assert np.array_equal(nd[ok], np.asarray(nd_ref).astype(np.uint32)[ok]) or int(flips.sum()) <= max(1, int(FLIPS_PER_NEUTRON * n))
    else:
        assert nbad == 0, "%s: %d of %d replayed neutrons differ by more than 1e-10" % (what, nbad, n)
    return nbad
