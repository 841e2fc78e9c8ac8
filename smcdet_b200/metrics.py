"""Catalog matching metrics with the reference's interface (smcdet/metrics.py:8-92).

``match_catalogs`` solves one small rectangular assignment problem per (tile, drawn posterior catalog); the
reference loops over them in Python and calls ``scipy.optimize.linear_sum_assignment`` each time
(metrics.py:36-61).  Here all T x n problems are one launch of ``smcdet_match_catalogs`` (one thread per
problem, scipy's algorithm restated in float64 so the assignments are the same).
"""

import torch

from . import _lib as L


def convert_nmgy_to_mag(nmgy):
    """reference utils/sdss.py:8-9"""
    return 22.5 - 2.5 * torch.log10(nmgy)


def match_catalogs(true_counts, true_locs, true_fluxes, est_counts, est_locs, est_fluxes, num_est_catalogs_to_match,
                   locs_tol, mags_tol, mag_bins, *, index=None):
    """Four [num_tiles, num_est_catalogs_to_match, len(mag_bins)] tensors: true stars per magnitude bin, matched
    true stars, estimated stars, matched estimated stars (reference metrics.py:8-84).  ``index`` (keyword-only,
    [num_tiles, num_est_catalogs_to_match]) fixes which estimated catalogs are matched; by default they are
    drawn uniformly as the reference does (metrics.py:40)."""
    dev = L.device()
    tc = L.f32(true_counts, dev).reshape(-1)
    T = tc.shape[0]
    tl, tf = L.f32(true_locs, dev).reshape(T, -1, 2), L.f32(true_fluxes, dev).reshape(T, -1)
    ec = L.f32(est_counts, dev).reshape(T, -1)
    M = ec.shape[1]
    ef = L.f32(est_fluxes, dev).reshape(T, M, -1)
    el = L.f32(est_locs, dev).reshape(T, M, -1, 2)
    n = int(num_est_catalogs_to_match)
    if index is None:
        index = torch.randint(0, M, (T, n), device=dev)
    index = index.to(device=dev, dtype=torch.int64).reshape(T, n).contiguous()
    if int(index.min()) < 0 or int(index.max()) >= M:
        raise IndexError("catalog index out of range")
    bins = L.f32(torch.as_tensor(mag_bins), dev).reshape(-1)
    B = bins.shape[0]
    out = [torch.zeros(T, n, B, device=dev) for _ in range(4)]
    status = torch.zeros(1, device=dev, dtype=torch.int32)
    L.check(L.lib().smcdet_match_catalogs(L.ptr(tc), L.ptr(tl), L.ptr(tf), L.ptr(ec), L.ptr(el), L.ptr(ef),
                                          L.ptr(index, torch.int64), L.ptr(bins), float(locs_tol), float(mags_tol),
                                          *[L.ptr(o) for o in out], L.ptr(status, torch.int32), T, n, M, tf.shape[1],
                                          ef.shape[2], B, L.stream_for(tc)))
    if int(status.item()) != 0:
        raise ValueError("match_catalogs: a catalog has more stars than its tensor holds or than the kernel supports (96)")
    return tuple(out)


def compute_precision_recall_f1(true_total, true_matches, est_total, est_matches):
    """reference metrics.py:87-92"""
    precision = (est_matches.sum(0) / est_total.sum(0)).nan_to_num(0)
    recall = (true_matches.sum(0) / true_total.sum(0)).nan_to_num(0)
    f1 = ((2 * precision * recall) / (precision + recall)).nan_to_num(0)
    return precision, recall, f1
