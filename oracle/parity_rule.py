"""ORACLE-SIDE TEST INFRASTRUCTURE: the acceptance rule of the fp32 CUDA path against the fp64 oracle.

north_star tolerances: log psi^2 1e-5, E_loc 1e-4 (relative).  A Slater matrix with condition number kappa amplifies the
fp32 rounding of its entries by kappa, so no fp32 evaluation -- the reference's own fp32 jax path included -- can hold a fixed
tolerance for every walker of a random-init network.  The rule therefore has two parts, both of which can fail:

  * walkers with cond_eff < COND_OK (1e3) must meet the stated tolerance OUTRIGHT (each of them);
  * every other walker is held to FLOOR_FACTOR (2) x the fp32 floor AT ITS CONDITIONING, where the floor is the worst error the
    fp32 CPU restatement (same algorithm, fp32 arithmetic, `floor`) makes on the same batch among walkers of the same or a lower
    cond_eff decade (the error grows with kappa, so the walkers of the same decade dominate that maximum), and never below the
    stated tolerance.

cond_eff (oracle/model.py::forward_laplacian, key "cond") = sum_d |q_d| cond_2(A_d) / |sum_d q_d|.
`table()` renders the error-vs-conditioning table that profiles/ keeps."""
from __future__ import annotations

import numpy as np

COND_OK = 1e3
FLOOR_FACTOR = 2.0


def _np(x):
    return np.asarray(x.detach().cpu() if hasattr(x, "detach") else x, dtype=np.float64)


def bounds(floor, cond, tol):
    """Per-walker bound of the rule."""
    floor, cond = _np(floor), _np(cond)
    dec = np.ceil(np.log10(np.maximum(cond, 1.0)))
    out = np.full_like(floor, tol)
    for i in range(len(floor)):
        if cond[i] >= COND_OK:
            out[i] = max(tol, FLOOR_FACTOR * floor[dec <= dec[i]].max())
    return out


def check(err, floor, cond, tol, what=""):
    """Raises AssertionError naming the offending walkers; returns the number of walkers judged at the plain tolerance."""
    err, cond = _np(err), _np(cond)
    b = bounds(floor, cond, tol)
    bad = np.nonzero(~(err <= b))[0]
    if len(bad):
        rows = ", ".join(f"walker {i}: err {err[i]:.2e} > bound {b[i]:.2e} (cond {cond[i]:.1e}, fp32 floor {_np(floor)[i]:.2e})" for i in bad[:6])
        raise AssertionError(f"{what}: {len(bad)} of {len(err)} walkers outside the parity rule (tol {tol:g}): {rows}")
    return int((cond < COND_OK).sum())


def table(err, floor, cond, label=""):
    """Markdown rows: cond decade | walkers | CUDA median / max | fp32-CPU median / max."""
    err, floor, cond = _np(err), _np(floor), _np(cond)
    dec = np.ceil(np.log10(np.maximum(cond, 1.0))).astype(int)
    lines = []
    for d in sorted(set(dec.tolist())):
        m = dec == d
        lines.append(f"| {label} | <1e{d} | {int(m.sum())} | {np.median(err[m]):.2e} | {err[m].max():.2e} | {np.median(floor[m]):.2e} | {floor[m].max():.2e} |")
    return lines
