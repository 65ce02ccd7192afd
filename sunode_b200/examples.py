"""The benchmark problems of BASELINE.json as ``SympyProblem`` definitions + their synthetic
inputs: Lotka-Volterra (reference README.md:56-118), Robertson, and a
two-group SEIR model.  Used by ``bench.py``, ``__graft_entry__`` and the tests; the problems are
ordinary user-level definitions, nothing here is special-cased by the engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Tuple

import numpy as np

from .symode import SympyProblem


def lotka_volterra() -> SympyProblem:
    """README.md:57-93 of the reference; derivatives wrt alpha and beta (README.md:85-90)."""
    def rhs(t, y, p):
        return {
            'hares': p.alpha * y.hares - p.beta * y.lynx * y.hares,
            'lynx': p.delta * y.hares * y.lynx - p.gamma * y.lynx,
        }
    return SympyProblem(
        params={'alpha': (), 'beta': (), 'gamma': (), 'delta': ()},
        states={'hares': (), 'lynx': ()},
        rhs_sympy=rhs,
        derivative_params=[('alpha',), ('beta',)],
    )


def robertson() -> SympyProblem:
    def rhs(t, y, p):
        return {
            'y1': -p.k1 * y.y1 + p.k3 * y.y2 * y.y3,
            'y2': p.k1 * y.y1 - p.k3 * y.y2 * y.y3 - p.k2 * y.y2 ** 2,
            'y3': p.k2 * y.y2 ** 2,
        }
    return SympyProblem(
        params={'k1': (), 'k2': (), 'k3': ()},
        states={'y1': (), 'y2': (), 'y3': ()},
        rhs_sympy=rhs,
        derivative_params=[('k1',), ('k2',), ('k3',)],
    )


def seir() -> SympyProblem:
    """Two-group SEIR, states (S, E, I, R) x 2 as population fractions, 6 parameters."""
    def rhs(t, y, p):
        out = {}
        betas = (p.beta1, p.beta2)
        groups = (y.g1, y.g2)
        for k, name in enumerate(('g1', 'g2')):
            g, h = groups[k], groups[1 - k]
            lam = betas[k] * g.I + p.kappa * h.I
            out[name] = {
                'S': -g.S * lam + p.omega * g.R,
                'E': g.S * lam - p.sigma * g.E,
                'I': p.sigma * g.E - p.gamma * g.I,
                'R': p.gamma * g.I - p.omega * g.R,
            }
        return out
    group = {'S': (), 'E': (), 'I': (), 'R': ()}
    names = ['beta1', 'beta2', 'kappa', 'sigma', 'gamma', 'omega']
    return SympyProblem(
        params={n: () for n in names},
        states={'g1': dict(group), 'g2': dict(group)},
        rhs_sympy=rhs,
        derivative_params=[(n,) for n in names],
    )


@dataclass
class Workload:
    name: str
    make_problem: Callable[[], SympyProblem]
    theta_med: Tuple[float, ...]
    sigma: float
    y0: Tuple[float, ...]
    t0: float
    tvals: np.ndarray
    batch: int
    seed: int
    adjoint: bool
    history_capacity: int
    cotangent: str = 'ones'
    sens: bool = False          # forward sensitivities dy/dp (Solver(sens_mode=...)) instead of the adjoint

    def grads(self, n_states: int) -> np.ndarray:
        """Cotangent ``g[n_t, n_s]`` shared by all instances: all ones as in the reference's
        smoke test (sunode/test_solve.py:99), or seeded N(0, 1) for the conservative systems
        (Robertson, SEIR), where ``sum_i y_i`` is constant so the all-ones cotangent has an
        identically zero gradient and a trivial backward problem."""
        if self.cotangent == 'ones':
            return np.ones((len(self.tvals), n_states))
        rng = np.random.default_rng(self.seed + 1000)
        return rng.standard_normal((len(self.tvals), n_states))

    def draws(self, batch: int = None, offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """``(y0[B, n_s], params[B, n_all])``: theta = theta_med * exp(sigma * N(0, 1)), i.i.d.
        per instance and parameter; the full-batch draw is generated and sliced so that shards
        of a multi-GPU run see exactly the instances a single-GPU run would."""
        batch = self.batch if batch is None else batch
        rng = np.random.default_rng(self.seed)
        z = rng.standard_normal((max(self.batch, offset + batch), len(self.theta_med)))
        theta = np.asarray(self.theta_med) * np.exp(self.sigma * z[offset:offset + batch])
        y0 = np.broadcast_to(np.asarray(self.y0, dtype=np.float64), (batch, len(self.y0)))
        return np.ascontiguousarray(y0), np.ascontiguousarray(theta)


def workloads() -> Dict[str, Workload]:
    """The BASELINE.json configs (cfg ids as in SURVEY.md 8d)."""
    lv_t = np.linspace(0, 10)
    return {
        'lv_fwd': Workload('lv_fwd', lotka_volterra, (0.1, 0.2, 0.3, 0.4), 0.25, (1.0, 0.1), 0.0,
                           lv_t, 65536, 20261017 + 2, False, 512),
        # forward sensitivity analysis of the same problem (dy/dalpha, dy/dbeta)
        'lv_fsa': Workload('lv_fsa', lotka_volterra, (0.1, 0.2, 0.3, 0.4), 0.25, (1.0, 0.1), 0.0,
                           lv_t, 65536, 20261017 + 2, False, 512, sens=True),
        'lv_adj': Workload('lv_adj', lotka_volterra, (0.1, 0.2, 0.3, 0.4), 0.25, (1.0, 0.1), 0.0,
                           lv_t, 65536, 20261017 + 2, True, 512),
        'robertson_adj': Workload('robertson_adj', robertson, (0.04, 3e7, 1e4), 0.1,
                                  (1.0, 0.0, 0.0), 0.0, np.logspace(-4, 4, 50), 16384,
                                  20261017 + 4, True, 4096, 'normal'),
        'seir_adj': Workload('seir_adj', seir, (0.5, 0.3, 0.05, 0.2, 0.1, 0.01), 0.2,
                             (0.99, 0.0, 0.01, 0.0, 0.995, 0.0, 0.005, 0.0), 0.0,
                             np.linspace(2, 100, 50), 262144, 20261017 + 5, True, 512, 'normal'),
    }


def problem_list() -> List[Callable[[], SympyProblem]]:
    return [lotka_volterra, robertson, seir]
