"""kmerml.ml.clustering -- the entry points the reference only declares (kmerml/ml/clustering.py:6-16 are `pass`
stubs), given bodies on top of the GPU distance matrix: the genome x genome distances come from
`KmerFeatureBuilder.distance_matrix` / `engine.pairwise_distance_device` (exact Gram matrix on the tensor cores for
count rows); the clustering itself is a few hundred points and runs in scipy / scikit-learn on the host.

Every function takes the organisms x k-mers feature matrix (DataFrame / array) -- the reference's signature -- and,
keyword-only, `distances=` (a precomputed square matrix, e.g. from another rank) and `metric=`."""
import numpy as np


def _distances(feature_matrix, distances, metric):
    if distances is not None:
        d = np.asarray(getattr(distances, "to_numpy", lambda: distances)(), dtype=np.float64)
        if d.ndim != 2 or d.shape[0] != d.shape[1]:
            raise ValueError("distances must be a square matrix")
        return d
    import torch
    from .. import engine
    x = np.ascontiguousarray(getattr(feature_matrix, "to_numpy", lambda: feature_matrix)())
    if np.issubdtype(x.dtype, np.integer) and x.min(initial=0) >= 0 and x.max(initial=0) < 2 ** 32:
        t = torch.from_numpy(x.astype(np.uint32).view(np.int32)).cuda()        # count rows: exact integer Gram
        if metric != "cosine":                                                  # Euclidean on counts as they are
            t = torch.from_numpy(x.astype(np.float64)).cuda()
    else:
        t = torch.from_numpy(x.astype(np.float64)).cuda()
    return engine.pairwise_distance_device(t, metric, out_dtype=torch.float64).cpu().numpy()


def hierarchical_clustering(feature_matrix, n_clusters=None, method="ward", *, distances=None, metric="euclidean"):
    """Agglomerative clustering of the organisms.  Returns (labels or None, linkage matrix): labels 0..n_clusters-1
    when n_clusters is given.  `method` as in scipy.cluster.hierarchy.linkage ("ward" expects Euclidean distances)."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform
    d = _distances(feature_matrix, distances, metric)
    d = (d + d.T) / 2.0
    np.fill_diagonal(d, 0.0)
    z = linkage(squareform(d, checks=False), method=method)
    labels = None if n_clusters is None else fcluster(z, t=int(n_clusters), criterion="maxclust") - 1
    return labels, z


def kmeans_clustering(feature_matrix, n_clusters=3, random_state=42):
    """k-means on the feature rows (scikit-learn); returns (labels, cluster centres)."""
    from sklearn.cluster import KMeans
    x = np.asarray(getattr(feature_matrix, "to_numpy", lambda: feature_matrix)(), dtype=np.float64)
    km = KMeans(n_clusters=int(n_clusters), random_state=random_state, n_init=10).fit(x)
    return km.labels_, km.cluster_centers_


def dbscan_clustering(feature_matrix, eps=0.5, min_samples=5, *, distances=None, metric="cosine"):
    """DBSCAN on the precomputed genome x genome distances; returns labels (-1 = noise)."""
    from sklearn.cluster import DBSCAN
    d = _distances(feature_matrix, distances, metric)
    return DBSCAN(eps=eps, min_samples=min_samples, metric="precomputed").fit(np.maximum(d, 0.0)).labels_
