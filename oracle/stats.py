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


# ---- clustering-coefficient and spectral MMD (evaluation/stats.py:61-148; kernels as utils/loader.py:201-206 selects
# them: gaussian_emd for both) on dense thresholded adjacency, no networkx -------------------------------------------
def emd1d_scaled(x, y, distance_scaling=1.0):
    """evaluation/mmd.py:16-29: ground distance |i - j| / distance_scaling."""
    return emd1d(x, y) / distance_scaling


def gaussian_emd_scaled(x, y, sigma=1.0, distance_scaling=1.0):
    e = emd1d_scaled(x, y, distance_scaling)
    return np.exp(-e * e / (2 * sigma * sigma))


def clustering_histograms(adj_q, counts, bins=100):
    """nx.clustering per node (triangles through the node / (deg (deg - 1) / 2), 0 when deg < 2), histogram over [0, 1] with
    `bins` bins (evaluation/stats.py:104-109)."""
    out = []
    for a, n in zip(np.asarray(adj_q), counts):
        a = a[:n, :n].astype(np.float64)
        deg = a.sum(axis=1)
        tri = np.einsum("ij,jk,ki->i", a, a, a) / 2.0
        denom = deg * (deg - 1) / 2.0
        cc = np.where(denom > 0, tri / np.maximum(denom, 1e-30), 0.0)
        h, _ = np.histogram(cc, bins=bins, range=(0.0, 1.0), density=False)
        out.append(h)
    return out


def spectral_pmfs(adj_q, counts):
    """eigenvalues of the normalised Laplacian I - D^-1/2 A D^-1/2 (isolated nodes: zero row, as networkx builds it),
    histogram with 200 bins over (-1e-5, 2), normalised (evaluation/stats.py:61-65)."""
    out = []
    for a, n in zip(np.asarray(adj_q), counts):
        if n == 0:
            continue
        a = a[:n, :n].astype(np.float64)
        deg = a.sum(axis=1)
        dis = np.where(deg > 0, 1.0 / np.sqrt(np.maximum(deg, 1e-30)), 0.0)
        lap = np.diag((deg > 0).astype(np.float64)) - dis[:, None] * a * dis[None, :]
        eig = np.linalg.eigvalsh(lap)
        h, _ = np.histogram(eig, bins=200, range=(-1e-5, 2), density=False)
        out.append(h / h.sum())
    return out


def mmd(hists_ref, hists_pred, kernel):
    """evaluation/mmd.py:91-100 `compute_mmd` (histograms normalised to pmfs)."""
    s1 = [np.asarray(h, dtype=np.float64) / np.sum(h) for h in hists_ref if np.sum(h) > 0]
    s2 = [np.asarray(h, dtype=np.float64) / np.sum(h) for h in hists_pred if np.sum(h) > 0]
    return disc(s1, s1, kernel) + disc(s2, s2, kernel) - 2 * disc(s1, s2, kernel)


def clustering_mmd(adj_ref, counts_ref, adj_pred, counts_pred, bins=100):
    """evaluation/stats.py:113-148 with KERNEL = gaussian_emd: sigma = 1 / 10, distance_scaling = bins."""
    k = lambda x, y: gaussian_emd_scaled(x, y, sigma=0.1, distance_scaling=bins)
    return mmd(clustering_histograms(adj_ref, counts_ref, bins), clustering_histograms(adj_pred, counts_pred, bins), k)


def spectral_mmd(adj_ref, counts_ref, adj_pred, counts_pred):
    """evaluation/stats.py:69-101 with KERNEL = gaussian_emd (sigma = 1, distance_scaling = 1)."""
    return mmd(spectral_pmfs(adj_ref, counts_ref), spectral_pmfs(adj_pred, counts_pred), gaussian_emd)
