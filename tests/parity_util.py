"""Strict parity gate for chains whose arithmetic goes through transcendental functions (device exp/log vs glibc).

A GPU chain either tracks the oracle to the contract tolerance on every draw, or it leaves the oracle's path at a draw
where the ORACLE'S OWN accept test was decided by a rounding-level margin |u - exp(comp)| — the only place where a
last-bit difference of exp/log may legitimately change the outcome.  Everything before that draw must match to the
tolerance, and the draw itself must be a flipped decision (one side kept its previous state).
One more legitimate cause exists for the reference's RM-HMC: it ACCEPTS non-finite energies and its fixed-point iterations
diverge for large steps, so some chains are numerically unstable — the ORACLE ITSELF, restarted from an initial point moved
by one ulp, leaves its own path by more than the tolerance.  Such a chain carries no parity information from that draw on;
it is recognised by exactly that experiment (`rerun_perturbed`) and reported, never silently skipped.  Anything else fails."""
import numpy as np

FLIP_MARGIN = 1e-9


def close_nan(a, b, tol):
    """L-inf agreement where both are finite, identical NaN pattern elsewhere (the reference ACCEPTS a NaN energy,
    src/rmhmc.cpp:250, so an unstable chain turns NaN at the same draw in the reference, the oracle and the kernel)."""
    return bool(np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a, b, rtol=0, atol=tol, equal_nan=True))


def assert_tracks_or_flips_at_threshold(gpu_draws, oracle_out, n_burnin, tol, what="", rerun_perturbed=None):
    """gpu_draws: [n_keep][d] of one chain; oracle_out: Oracle.run_chain(..., want_margins=True) of the same chain;
    rerun_perturbed: optional callable returning the oracle's draws of the same chain from a start moved by one ulp.
    Returns True if the chain tracked on every draw, False if it flipped at a rounding-level margin or is numerically
    unstable in the oracle itself from that draw on (asserts otherwise)."""
    od = oracle_out["draws"]
    n_keep = od.shape[0]
    first_bad = None
    for t in range(n_keep):
        if not close_nan(gpu_draws[t], od[t], tol):
            first_bad = t
            break
    if first_bad is None:
        return True
    # a proposal that overflowed (|x| beyond 1e30 or non-finite on either side): what the accept test then sees is inf / NaN
    # arithmetic whose propagation through an LU, a Cholesky factor and a quadratic form is implementation-defined
    blown = lambda a: (not np.isfinite(a).all()) or np.abs(a[np.isfinite(a)]).max(initial=0.0) > 1e30
    if blown(gpu_draws[first_bad:]) or blown(od[first_bad:]):
        print("%s: overflowed trajectory from kept draw %d on (non-finite arithmetic in the accept test)" % (what, first_bad))
        return False
    if rerun_perturbed is not None:
        pd = rerun_perturbed()
        if not close_nan(pd[first_bad], od[first_bad], 0.25 * tol):
            print("%s: numerically unstable in the oracle itself at kept draw %d (one-ulp start perturbation moves it by %.2e)"
                  % (what, first_bad, float(np.nanmax(np.abs(pd[first_bad] - od[first_bad])))))
            return False
    if first_bad == 0 and n_burnin > 0:   # the flip may sit in the burn-in, whose draws are not returned
        m = oracle_out["margins"][:n_burnin + 1]
        m = m[np.argmin(np.abs(m))]
    else:
        m = oracle_out["margins"][n_burnin + first_bad]
    assert np.isfinite(m) and abs(m) <= FLIP_MARGIN, (
        "%s: chain leaves the oracle's path at kept draw %d where the accept margin u - exp(comp) = %.3e is not at rounding level"
        % (what, first_bad, m))
    prev_g = gpu_draws[first_bad - 1] if first_bad > 0 else None
    prev_o = od[first_bad - 1] if first_bad > 0 else None
    if prev_g is not None and not (first_bad == 0):
        kept_g = close_nan(gpu_draws[first_bad], prev_g, 0.0)
        kept_o = close_nan(od[first_bad], prev_o, 0.0)
        assert kept_g != kept_o, "%s: divergence at draw %d is not a flipped accept decision" % (what, first_bad)
    return False
