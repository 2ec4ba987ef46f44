"""fp64 companion of the codebook lookup, used to classify near-ties.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference computes ``d = (|z|^2 + |e|^2) - 2 z.e`` in fp32 with whatever
summation order its BLAS picks (model/codebook.py:19-21), so two correct fp32
evaluations may disagree on rows whose best and second-best distance are closer
than a few ulp of ``|z|^2``.  Parity is therefore stated as:

  * rows whose fp64 margin (second-best minus best distance) exceeds ``tau``
    must return exactly the fp64 arg-min;
  * on the remaining rows the chosen code's fp64 distance may exceed the fp64
    minimum by at most ``tau`` ("regret"), and exact ties go to the lowest index.

``tau`` defaults to 16 ulp of max(|z|^2 + |e|^2): the rounding a length-D fp32
accumulation can accrue.
"""
from __future__ import annotations

import numpy as np


def distances64(z_rows, emb):
    """z_rows: [N, D]; emb: [K, D] -> fp64 [N, K] squared distances."""
    z = np.asarray(z_rows, dtype=np.float64)
    e = np.asarray(emb, dtype=np.float64)
    return (z * z).sum(1)[:, None] + (e * e).sum(1)[None, :] - 2.0 * (z @ e.T)


def classify(z_rows, emb, idx, tau=None, chunk=8192):
    """Compare ``idx`` (int array [N]) with the fp64 ranking.

    Returns dict(n, exact, clear_rows, clear_mismatch, max_regret, tau).
    """
    z_rows = np.asarray(z_rows)
    emb = np.asarray(emb)
    idx = np.asarray(idx).astype(np.int64)
    n = z_rows.shape[0]
    exact = 0
    clear_rows = 0
    clear_mismatch = 0
    max_regret = 0.0
    tau_used = 0.0
    for s in range(0, n, chunk):
        d = distances64(z_rows[s:s + chunk], emb)
        if tau is None:
            scale = float(np.max((z_rows[s:s + chunk].astype(np.float64) ** 2).sum(1))
                          + np.max((emb.astype(np.float64) ** 2).sum(1)))
            t = 16.0 * np.spacing(np.float32(scale)).astype(np.float64)
        else:
            t = float(tau)
        tau_used = max(tau_used, t)
        best = d.argmin(1)
        part = np.partition(d, 1, axis=1)
        margin = part[:, 1] - part[:, 0]
        got = idx[s:s + chunk]
        rows = np.arange(d.shape[0])
        regret = d[rows, got] - d[rows, best]
        exact += int((got == best).sum())
        clear = margin > t
        clear_rows += int(clear.sum())
        clear_mismatch += int((clear & (got != best)).sum())
        max_regret = max(max_regret, float(regret.max(initial=0.0)))
    return dict(n=n, exact=exact, clear_rows=clear_rows, clear_mismatch=clear_mismatch,
                max_regret=max_regret, tau=tau_used)
