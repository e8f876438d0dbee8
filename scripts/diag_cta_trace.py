#!/usr/bin/env python
"""Diagnostic: read the per-CTA wall-clock marks k_frame_fused writes under HBGPU_CTA_TRACE=<file> (one record of
8 u64 per frame: start, after the stagger, B begin, B end, B2 end, end, SM id) and print how the frames of a launch
overlap: phase durations, how many CTAs stream at once (whole GPU / per SM) and how a CTA's streaming time depends on it.
usage: diag_cta_trace.py trace.bin frames_per_launch [launch index, default last]"""
import sys

import numpy as np


def main():
    path, nf = sys.argv[1], int(sys.argv[2])
    a = np.fromfile(path, dtype=np.uint64).reshape(-1, 8)
    nl = a.shape[0] // nf
    k = int(sys.argv[3]) if len(sys.argv) > 3 else nl - 1
    r = a[k * nf:(k + 1) * nf].astype(np.int64)
    t = r[:, :6] - r[:, 0].min()
    sm = r[:, 6]
    ok = (r[:, 5] > 0)
    print("launches in file %d, frames %d, complete records %d, SMs %d" % (nl, nf, ok.sum(), len(np.unique(sm))))
    t = t[ok]; sm = sm[ok]
    names = ["stagger", "A+g1", "B stream", "B2", "C"]
    d = np.diff(t, axis=1)
    print("kernel span %.1f us" % (t[:, 5].max() / 1e3))
    for i, n in enumerate(names):
        print("  %-9s mean %8.1f us  p10 %8.1f  p50 %8.1f  p90 %8.1f" % (n, d[:, i].mean() / 1e3, *(np.percentile(d[:, i], [10, 50, 90]) / 1e3)))
    life = t[:, 5] - t[:, 0]
    print("  frame     mean %8.1f us" % (life.mean() / 1e3))
    # first wave: which block ids share an SM
    first = np.argsort(r[ok][:, 0])[:16]
    order = np.argsort(t[:, 0], kind="stable")
    print("first-wave mapping (block -> SM) of blocks 0..11 and 148..151:", [(int(b), int(sm[b])) for b in list(range(12)) + list(range(148, 152)) if b < len(sm)])
    # concurrency over time
    grid = np.linspace(0, t[:, 5].max(), 2001)
    bs, be = t[:, 2], t[:, 3]
    stream_all = np.array([((bs <= g) & (be > g)).sum() for g in grid])
    alive_all = np.array([((t[:, 0] <= g) & (t[:, 5] > g)).sum() for g in grid])
    mid = slice(200, 1800)
    print("CTAs alive (steady) mean %.0f; streaming mean %.0f  min %d  max %d  std %.0f" % (
        alive_all[mid].mean(), stream_all[mid].mean(), stream_all[mid].min(), stream_all[mid].max(), stream_all[mid].std()))
    # autocorrelation-ish: show the streaming count along time (40 samples)
    print("streaming CTAs along the launch:", " ".join("%d" % v for v in stream_all[::50]))
    # per-CTA: average number of CTAs streaming on the same SM / on the GPU while it streams
    nsm = int(sm.max()) + 1
    per_sm_idx = [np.nonzero(sm == s)[0] for s in range(nsm)]
    same = np.zeros(len(t)); glob = np.zeros(len(t))
    for s in range(nsm):
        idx = per_sm_idx[s]
        for i in idx:
            ov = np.minimum(be[idx], be[i]) - np.maximum(bs[idx], bs[i])
            same[i] = np.clip(ov, 0, None).sum() / max(be[i] - bs[i], 1)
    samp = np.random.default_rng(0).choice(len(t), size=min(600, len(t)), replace=False)
    for i in samp:
        ov = np.minimum(be, be[i]) - np.maximum(bs, bs[i])
        glob[i] = np.clip(ov, 0, None).sum() / max(be[i] - bs[i], 1)
    bdur = (be - bs) / 1e3
    print("streaming CTAs on the same SM while a CTA streams: mean %.2f" % same.mean())
    for lo, hi in [(0, 1.5), (1.5, 2.5), (2.5, 3.5), (3.5, 9)]:
        m = (same >= lo) & (same < hi)
        if m.sum():
            print("  same-SM streams in [%.1f, %.1f): %5d CTAs, B mean %.1f us" % (lo, hi, m.sum(), bdur[m].mean()))
    g = glob[samp]
    qs = np.percentile(g, [0, 25, 50, 75, 100])
    for lo, hi in zip(qs[:-1], qs[1:]):
        m = (g >= lo) & (g <= hi)
        print("  GPU-wide streams in [%.0f, %.0f]: B mean %.1f us (same-SM %.2f)" % (lo, hi, bdur[samp][m].mean(), same[samp][m].mean()))
    print("corr(B duration, same-SM streams) %.2f   corr(B duration, GPU-wide streams) %.2f" % (
        np.corrcoef(bdur, same)[0, 1], np.corrcoef(bdur[samp], g)[0, 1]))


if __name__ == "__main__":
    main()
