"""Graph statistics for the large-batch statistical parity check (SURVEY.md 8c protocol iv) -- TEST INFRASTRUCTURE ONLY.

Restatement of the degree MMD of the reference's evaluation (evaluation/stats.py:28-58 `degree_stats`,
evaluation/mmd.py:16-29 `emd`, :37-45 `gaussian_emd`, :73-100 `disc` / `compute_mmd`) on thresholded dense adjacency
tensors instead of networkx graphs.  The reference computes the earth mover's distance with `pyemd` (not installed
here) on a Toeplitz ground distance |i - j|; for 1-D histograms with that ground distance the EMD is the L1 distance of
the two CDFs, which is what `emd1d` computes (unit-mass pmfs of equal length)."""
import numpy as np


def edge_counts(adj_q):
    """adj_q [B,N,N] in {0,1}, symmetric, zero diagonal -> edges per graph."""
    return np.asarray(adj_q).sum(axis=(1, 2)) / 2


def degree_histograms(adj_q, counts):
    """Per-graph degree histogram over the graph's own nodes (nx.degree_histogram, evaluation/stats.py:49): list of
    integer arrays of length max_degree + 1."""
    out = []
    for a, n in zip(np.asarray(adj_q), counts):
        deg = a[:n, :n].sum(axis=1).astype(int)
        out.append(np.bincount(deg))
    return out


def emd1d(x, y):
    """evaluation/mmd.py:16-29 with distance_scaling = 1: pad to equal support, EMD for the ground distance |i - j|."""
    m = max(len(x), len(y))
    x = np.pad(np.asarray(x, dtype=np.float64), (0, m - len(x)))
    y = np.pad(np.asarray(y, dtype=np.float64), (0, m - len(y)))
    return float(np.abs(np.cumsum(x - y)).sum())


def gaussian_emd(x, y, sigma=1.0):
    """evaluation/mmd.py:37-45."""
    e = emd1d(x, y)
    return np.exp(-e * e / (2 * sigma * sigma))


def disc(s1, s2, kernel):
    """evaluation/mmd.py:73-88."""
    return sum(kernel(a, b) for a in s1 for b in s2) / (len(s1) * len(s2))


def degree_mmd(hists_ref, hists_pred, kernel=gaussian_emd):
    """evaluation/mmd.py:91-100 `compute_mmd` on degree histograms (normalised to pmfs), as `degree_stats` calls it."""
    s1 = [h / h.sum() for h in hists_ref if h.sum() > 0]
    s2 = [h / h.sum() for h in hists_pred if h.sum() > 0]
    return disc(s1, s1, kernel) + disc(s2, s2, kernel) - 2 * disc(s1, s2, kernel)
