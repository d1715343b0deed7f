#!/usr/bin/env python3
"""Fit synth.COST_TABLE (the per-record cost model used to size contiguous shards) to measured per-stream cycles.

    python bench.py --balance reads --dump-streams gpurun_out/streams.npz ...      (on the GPU box)
    python bench/fit_cost_model.py gpurun_out/streams.npz                          (here)

Model: cost(record) = COST_DUP for a duplicate (RepeatAddRead), else TABLE[bucket(minCnt)][bucket(medianCnt)];
non-negative least squares on the per-stream sums, then cells with too few reads are replaced by their row median
and everything is clipped to [60, 1000] microseconds."""
import sys

import numpy as np
from scipy.optimize import nnls

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from trust4_b200 import synth  # noqa: E402


def main():
    d = np.load(sys.argv[1])
    flags, mc, med, off = d["flags"], d["min_cnt"].astype(np.int64), d["med"].astype(np.int64), d["off"]
    y = d["cycles"].astype(float) / 1.965e3        # microseconds at 1965 MHz
    ME, DE = np.array(synth.COST_MIN_EDGES), np.array(synth.COST_MED_EDGES)
    cell = np.searchsorted(ME, mc, side="left") * len(DE) + np.searchsorted(DE, med, side="left")
    nc = len(ME) * len(DE)
    dup = (flags & 1) != 0
    S = len(off) - 1
    X = np.zeros((S, 1 + nc))
    for j in range(S):
        lo, hi = off[j], off[j + 1]
        nd = ~dup[lo:hi]
        X[j, 0] = (~nd).sum()
        X[j, 1:] = np.bincount(cell[lo:hi][nd], minlength=nc)
    coef, _ = nnls(X, y)
    pred = X @ coef
    T = coef[1:].reshape(len(ME), len(DE))
    T2 = T.copy()
    for i in range(T.shape[0]):
        nz = T[i][(T[i] > 0) & (T[i] < 1200)]
        fill = np.median(nz) if len(nz) else 200
        T2[i][(T[i] <= 0) | (T[i] > 1200)] = fill
    T2 = np.clip(T2, 60, 1000)
    r = y / np.maximum(pred, 1)
    print("dup %.2f us" % coef[0])
    print("COST_TABLE =", repr(np.round(T2).astype(int).tolist()))
    print("corr %.3f   measured/predicted: max %.2f  p99 %.2f   streams: mean %.1f ms max %.1f ms" %
          (np.corrcoef(pred, y)[0, 1], r.max(), np.percentile(r, 99), y.mean() / 1e3, y.max() / 1e3))


if __name__ == "__main__":
    main()
