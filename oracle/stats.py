"""Graph statistics for the large-batch statistical parity check (SURVEY.md 8c protocol iv) -- TEST INFRASTRUCTURE ONLY.

Restatement of the degree MMD of the reference's evaluation (evaluation/stats.py:28-58 `degree_stats`,
evaluation/mmd.py:16-29 `emd`, :37-45 `gaussian_emd`, :73-100 `disc` / `compute_mmd`) on thresholded dense adjacency
tensors instead of networkx graphs.  The reference computes the earth mover's distance with `pyemd` (not installed
here) on a Toeplitz ground distance |i - j|; for 1-D histograms with that ground distance the EMD is the L1 distance of
the two CDFs, which is what `emd1d` computes (unit-mass pmfs of equal length).  Further down: clustering / spectral / orbit
MMD (pinned by tests/golden/c1_mmd_reference.npz) and the NSPDK MMD of molecule graphs (tests/golden/nspdk_reference.npz)."""
import numpy as np


def edge_counts(adj_q):
    """adj_q [B,N,N] in {0,1}, symmetric, zero diagonal -> edges per graph."""
    return np.asarray(adj_q).sum(axis=(1, 2)) / 2


def degree_histograms(adj_q, counts):
    """Per-graph degree histogram over the graph's own nodes (nx.degree_histogram, evaluation/stats.py:49): list of
    integer arrays of length max_degree + 1."""
    out = []
    for a, n in zip(adj_q, counts):             # (a 3-D array or a list of per-graph arrays)
        deg = np.asarray(a)[:n, :n].sum(axis=1).astype(int)
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
    for a, n in zip(adj_q, counts):
        a = np.asarray(a)[:n, :n].astype(np.float64)
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
    for a, n in zip(adj_q, counts):
        if n == 0:
            continue
        a = np.asarray(a)[:n, :n].astype(np.float64)
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


# ---- orbit statistics (evaluation/stats.py:172-193 `orbit_stats_all`): per-node counts of the 15 orbits of the 2- to 4-node
# graphlets as the reference's ORCA binary reports them (`orca node 4`), restated by brute force over the induced 3- and 4-node
# subgraphs (fine for the <= 20-node graphs of the C1 check; pinned to ORCA's own output by tests/golden/c1_mmd_reference.npz)
_COMBOS = {}


def _combos(n, k):
    key = (n, k)
    if key not in _COMBOS:
        from itertools import combinations
        _COMBOS[key] = np.array(list(combinations(range(n), k)), dtype=np.int64).reshape(-1, k)
    return _COMBOS[key]


def orbit_counts(a):
    """a [n,n] in {0,1}, symmetric, zero diagonal -> int64 [n,15].  Orbits: 0 edge; 1 / 2 end / middle of the induced 3-path;
    3 triangle; 4 / 5 end / inner node of the 4-path; 6 / 7 leaf / centre of the claw; 8 4-cycle; 9 / 10 / 11 pendant node /
    degree-2 triangle nodes / degree-3 node of the tailed triangle; 12 / 13 degree-2 / degree-3 nodes of the diamond; 14 K4."""
    a = (np.asarray(a) > 0).astype(np.int64)
    n = a.shape[0]
    out = np.zeros((n, 15), dtype=np.int64)
    out[:, 0] = a.sum(axis=1)
    if n >= 3:
        c = _combos(n, 3)
        sub = a[c[:, :, None], c[:, None, :]]                 # [M,3,3]
        deg = sub.sum(axis=2)
        edges = deg.sum(axis=1) // 2
        for m in np.nonzero(edges == 2)[0]:
            for q in range(3):
                out[c[m, q], 1 if deg[m, q] == 1 else 2] += 1
        for m in np.nonzero(edges == 3)[0]:
            out[c[m], 3] += 1
    if n >= 4:
        c = _combos(n, 4)
        sub = a[c[:, :, None], c[:, None, :]]                 # [M,4,4]
        deg = sub.sum(axis=2)
        edges = deg.sum(axis=1) // 2
        conn = (deg.min(axis=1) > 0) & (edges >= 3)
        # two disjoint edges (degrees 1,1,1,1) have 2 edges and are excluded by edges >= 3; 3 edges with every degree > 0 is a
        # path or a claw (a triangle + isolated node has a zero degree)
        dmax = deg.max(axis=1)
        table = {  # (edges, max degree) -> {degree in the graphlet: orbit}
            (3, 2): {1: 4, 2: 5}, (3, 3): {1: 6, 3: 7}, (4, 2): {2: 8}, (4, 3): {1: 9, 2: 10, 3: 11}, (5, 3): {2: 12, 3: 13},
            (6, 3): {3: 14}}
        for m in np.nonzero(conn)[0]:
            orb = table[(int(edges[m]), int(dmax[m]))]
            for q in range(4):
                out[c[m, q], orb[int(deg[m, q])]] += 1
    return out


def graph_from_adj(a):
    """utils/graph_utils.py:165-176 `adjs_to_graphs` on one thresholded adjacency: self loops and isolated nodes removed; a
    graph left without nodes becomes one isolated node.  Returns the dense adjacency of what is left."""
    a = (np.asarray(a) > 0).astype(np.int8)
    np.fill_diagonal(a, 0)
    keep = a.sum(axis=1) > 0
    if not keep.any():
        return np.zeros((1, 1), dtype=np.int8)
    return a[np.ix_(keep, keep)]


def orbit_features(adjs):
    """per graph: orbit counts summed over the nodes / number of nodes (evaluation/stats.py:181, :190)."""
    return [orbit_counts(a).sum(axis=0) / a.shape[0] for a in adjs]


def gaussian_l2(x, y, sigma=1.0):
    """evaluation/mmd.py:48-53 `gaussian` (vectors padded to equal length)."""
    m = max(len(x), len(y))
    x = np.pad(np.asarray(x, dtype=np.float64), (0, m - len(x)))
    y = np.pad(np.asarray(y, dtype=np.float64), (0, m - len(y)))
    d = np.linalg.norm(x - y, 2)
    return np.exp(-d * d / (2 * sigma * sigma))


def orbit_mmd(adjs_ref, adjs_pred):
    """evaluation/stats.py:172-193 with KERNEL = gaussian: compute_mmd(..., is_hist=False, sigma=30)."""
    s1, s2 = orbit_features(adjs_ref), orbit_features(adjs_pred)
    k = lambda x, y: gaussian_l2(x, y, sigma=30.0)
    return disc(s1, s1, k) + disc(s2, s2, k) - 2 * disc(s1, s2, k)


def eval_graph_lists(adjs_ref, adjs_pred):
    """evaluation/stats.py:218-230 `eval_graph_list` with the methods / kernels of utils/loader.py:199-206 (degree, cluster,
    spectral: gaussian_emd; orbit: gaussian) on lists of dense adjacencies whose isolated nodes are already removed
    (`graph_from_adj` for generated samples; training graphs as they are).  Values rounded to 6 digits as the reference does."""
    cr, cp = [a.shape[0] for a in adjs_ref], [a.shape[0] for a in adjs_pred]
    pad = lambda lst: lst                                   # the histogram helpers take per-graph arrays + counts
    res = {"degree": degree_mmd(degree_histograms(pad(adjs_ref), cr), degree_histograms(pad(adjs_pred), cp)),
           "cluster": clustering_mmd(adjs_ref, cr, adjs_pred, cp), "orbit": orbit_mmd(adjs_ref, adjs_pred),
           "spectral": spectral_mmd(adjs_ref, cr, adjs_pred, cp)}
    return {k: round(float(v), 6) for k, v in res.items()}


# ---- NSPDK MMD (evaluation/stats.py:231-241 `nspdk_stats`, evaluation/mmd.py:115-127 `compute_nspdk_mmd`, evaluation/eden.py
# `vectorize(graphs, complexity=4, discrete=True)`): the neighbourhood-subgraph pairwise-distance kernel of EDeN as the
# reference vendors it, restated on (labels [n], bond orders [n,n]) arrays instead of networkx graphs.  The reference hashes
# with Python's builtin `hash` of int tuples (deterministic) and of the LABEL itself (evaluation/eden.py:943-948).  sampler.py
# labels atoms with their symbol string (utils/mol_utils.py:161-183), whose hash changes from process to process; here a label
# is an integer (the atomic number; bond order for an edge), for which hash(v) = v.  The kernel value of two graphs does not
# depend on the label -> code map except through collisions in the 2^16-entry feature space, so the restatement is pinned
# EXACTLY against the reference run on integer-labelled graphs and to the collision noise against the string-labelled run
# (tests/golden/nspdk_reference.npz).
_NSPDK_BITS = 16                                     # Vectorizer(nbits=16): evaluation/eden.py:160, :245-246
_NSPDK_MASK = (1 << _NSPDK_BITS) - 1                 # labels and feature ids
_NSPDK_MASK32 = 4294967295                           # evaluation/eden.py:22: the default mask of the neighbourhood hashes


def _h(mask, *items):
    """fast_hash_2 / _3 / _4 / fast_hash (evaluation/eden.py:57-70)."""
    return int(hash(tuple(items)) & mask) + 1


def _expanded(labels, bonds):
    """evaluation/eden.py:951-977 `_edge_to_vertex_transform` + :943-948 `_label_preprocessing`: one extra vertex per edge
    (self loops dropped), joined to its two endpoints.  -> (hashed label per vertex, neighbour lists, number of atom vertices)."""
    labels = [int(v) for v in labels]
    b = np.asarray(bonds)
    n = len(labels)
    hlabel = [int(hash(v) & _NSPDK_MASK) + 1 for v in labels]
    nbr = [[] for _ in range(n)]
    for u in range(n):
        for v in range(u + 1, n):
            if b[u, v] != 0:
                w = len(nbr)
                nbr.append([u, v])
                nbr[u].append(w)
                nbr[v].append(w)
                hlabel.append(int(hash(int(b[u, v])) & _NSPDK_MASK) + 1)
    return hlabel, nbr, n


def _levels(nbr, root, max_depth):
    """evaluation/eden.py:730-763: breadth-first levels of the expanded graph around `root`, up to `max_depth`."""
    seen = {root}
    levels = [[root]]
    while len(levels) <= max_depth:
        nxt = []
        for u in levels[-1]:
            for v in nbr[u]:
                if v not in seen:
                    seen.add(v)
                    nxt.append(v)
        if not nxt:
            break
        levels.append(nxt)
    return levels


def nspdk_features(labels, bonds, complexity=4):
    """One graph -> {feature id: value}, both normalisations on (evaluation/eden.py:421-430 `_transform`, :438-453, :547-616,
    :618-650).  Vertices of the expanded graph alternate atom / bond, so distance and radius step by 2."""
    hlabel, nbr, n = _expanded(labels, bonds)
    if n == 0:
        raise ValueError("empty graph (evaluation/eden.py:368-370)")
    r = d = complexity
    levels, nhash = [], []
    for root in range(n):
        lv = _levels(nbr, root, 2 * max(r, d))
        levels.append(lv)
        # evaluation/eden.py:658-689: per level the sorted multiset of hash(label, degree), hashed; then the running hash of
        # the level sequence (fast_hash_vec :73-79) -> one code per radius
        codes, run = [], 0xAAAAAAAA
        for k, level in enumerate(lv):
            item = _h(_NSPDK_MASK32, *sorted(_h(_NSPDK_MASK32, hlabel[v], len(nbr[v])) for v in level))
            run ^= hash((run, item, k))
            codes.append(int(run & _NSPDK_MASK32) + 1)
        nhash.append(codes)
    blocks = {}                                          # (radius / 2, distance / 2) -> {feature: count}, in first-use order
    for v in range(n):
        for dist in range(0, 2 * (d + 1), 2):
            if dist >= len(levels[v]):
                continue
            for u in levels[v][dist]:
                for rad in range(0, 2 * (r + 1), 2):
                    if rad < len(nhash[v]) and rad < len(nhash[u]):
                        hv, hu = nhash[v][rad], nhash[u][rad]
                        blk = blocks.setdefault((rad // 2, dist // 2), {})
                        full = _h(_NSPDK_MASK, min(hv, hu), max(hv, hu), rad, dist)
                        blk[full] = blk.get(full, 0.0) + 1.0
                        half = _h(_NSPDK_MASK, hu, rad, dist)                    # the context feature that ignores the centre vertex
                        blk[half] = blk.get(half, 0.0) + 1.0
    vec = {}
    for blk in blocks.values():                          # a feature id met again in a later block REPLACES the earlier value
        norm = np.sqrt(sum(c * c for c in blk.values()))
        for f, c in blk.items():
            vec[f] = float(c) / norm
    total = np.sqrt(sum(v * v for v in vec.values()))
    return {f: v / total for f, v in vec.items()}


def nspdk_matrix(graphs, complexity=4):
    """graphs: list of (labels [n], bonds [n,n]) -> sparse [G, 2^16 + 1] feature matrix (evaluation/eden.py:372-390)."""
    from scipy.sparse import csr_matrix
    data, row, col = [], [], []
    for i, (labels, bonds) in enumerate(graphs):
        for f, v in nspdk_features(labels, bonds, complexity).items():
            row.append(i)
            col.append(f)
            data.append(v)
    return csr_matrix((data, (row, col)), shape=(len(graphs), _NSPDK_MASK + 2), dtype=np.float64)


def nspdk_mmd(graphs_ref, graphs_pred, with_unbiased=False):
    """evaluation/mmd.py:115-127 with the linear kernel on the NSPDK vectors; graphs without nodes are dropped from the
    generated list as evaluation/stats.py:232 does.  The reference's number averages the kernel matrices WITH their diagonals
    (every vector has unit norm), so two lists drawn from one distribution still give about (1 / m + 1 / n) (1 - mean kernel).
    with_unbiased: also return the same statistic without the diagonals, which is 0 in expectation for one distribution --
    what a sample-against-sample parity check can put a threshold on."""
    x = nspdk_matrix(graphs_ref)
    y = nspdk_matrix([g for g in graphs_pred if len(g[0]) > 0])
    kxx, kyy, kxy = (x @ x.T).toarray(), (y @ y.T).toarray(), (x @ y.T).toarray()
    val = float(np.mean(kxx) + np.mean(kyy) - 2 * np.mean(kxy))
    if not with_unbiased:
        return val
    m, n = kxx.shape[0], kyy.shape[0]
    ub = (kxx.sum() - np.trace(kxx)) / (m * (m - 1)) + (kyy.sum() - np.trace(kyy)) / (n * (n - 1)) - 2 * np.mean(kxy)
    return val, float(ub)


QM9_ATOMS = (6, 7, 8, 9)                                 # utils/mol_utils.py:56 / :58 without the virtual-atom entry
ZINC_ATOMS = (6, 7, 8, 9, 15, 16, 17, 35, 53)


def mol_graphs(atoms, bonds, atomic_numbers):
    """atoms [B,N,F] in {0,1} (x > 0.5), bonds [B,N,N] bond orders 0..3 (quantize_mol) -> list of (atomic numbers [n], bond
    orders [n,n]) the way utils/mol_utils.py:72-90 `construct_mol` reads them: atom type = argmax over [x, 1 - sum x] (the last
    entry = no atom: dropped), bonds among the atoms that exist.  RDKit's valency correction / largest-fragment step that
    follows in gen_mol is NOT applied (RDKit is absent here)."""
    atoms, bonds = np.asarray(atoms).astype(np.int64), np.asarray(bonds).astype(np.int64)
    out = []
    for a, b in zip(atoms, bonds):
        t = np.argmax(np.concatenate([a, 1 - a.sum(axis=1, keepdims=True)], axis=1), axis=1)
        keep = t != a.shape[1]
        out.append((np.asarray(atomic_numbers)[t[keep]], b[np.ix_(keep, keep)]))
    return out
