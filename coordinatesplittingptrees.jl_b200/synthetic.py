"""Seeded synthetic trees for the benchmark / parity configurations of SURVEY.md 8(d).

Every tree is grown with the 4-argument addpoint! and explicit dimlists: CS1 trees use [(1,),(2,),...];
CS2 trees cycle through the n-1 perfect matchings of a round-robin 1-factorisation of K_n, so every pair (i,j) recurs
every n-1 points.  Points are uniform in the world box.
"""
import numpy as np

from . import tree as T


def round_robin_dimlists(n):
    """The n-1 (n even) or n (n odd) rounds of the circle method; each round is a list of pairs covering 1..n once
    (for odd n one dimension per round is left alone and listed last as a 1-tuple)."""
    m = n if n % 2 == 0 else n + 1
    ring = list(range(1, m + 1))
    rounds = []
    for _ in range(m - 1):
        pairs, single = [], None
        for i in range(m // 2):
            a, b = ring[i], ring[m - 1 - i]
            if a > n or b > n:
                single = (min(a, b),)
            else:
                pairs.append((a, b))
        rounds.append(pairs + ([single] if single else []))
        ring = [ring[0]] + [ring[-1]] + ring[1:-1]
    return rounds


def cs1_dimlists(n):
    return [[(d,) for d in range(1, n + 1)]]


def grow(n, p, npoints, objective, seed, lower=-1.0, upper=1.0, base=0.0, dimlist_cycle=None):
    """Root of a tree grown by `npoints` addpoint! calls at U(world) points."""
    rng = np.random.default_rng(seed)
    lo, up = np.full(n, float(lower)), np.full(n, float(upper))
    pos = np.full(n, float(base)) if np.isscalar(base) else np.asarray(base, np.float64)
    world = T.World(lo, up, pos, objective)
    root = T.Box(world, p)
    if dimlist_cycle is None:
        dimlist_cycle = cs1_dimlists(n) if p == 1 else round_robin_dimlists(n)
    X = lo + (up - lo) * rng.random((npoints, n))
    T.addpoints(root, X, dimlist_cycle, objective)
    return root


def quad_objective(n, seed, sigma=1e-3):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n))
    return T.quadnoise_objective((G + G.T) / 2, sigma=sigma, seed=seed)


CONFIGS = {
    # name: (p, n, addpoints, objective factory, world)  -- leaf counts as in SURVEY 8(d)
    "C1": dict(p=1, n=2, npoints=500, objective=lambda: T.rosenbrock_objective(), lower=-2.0, upper=2.0, base=(-1.2, 1.0)),
    "C2": dict(p=2, n=10, npoints=6667, objective=lambda: quad_objective(10, 2)),
    "C3": dict(p=2, n=20, npoints=33334, objective=lambda: quad_objective(20, 3)),
    "C4": dict(p=1, n=100, npoints=100, objective=lambda: quad_objective(100, 4, sigma=0.0)),
    "C5": dict(p=2, n=300, npoints=23, objective=lambda: quad_objective(300, 5, sigma=0.0)),
}


def config_tree(name, seed=0, npoints=None):
    c = CONFIGS[name]
    obj = c["objective"]()
    root = grow(c["n"], c["p"], npoints if npoints is not None else c["npoints"], obj, seed=1000 + seed,
                lower=c.get("lower", -1.0), upper=c.get("upper", 1.0), base=c.get("base", 0.0))
    return root, obj


def uniform_queries(root, m, seed):
    rng = np.random.default_rng(seed)
    w = root.world
    return w.lower + (w.upper - w.lower) * rng.random((m, root.n))
