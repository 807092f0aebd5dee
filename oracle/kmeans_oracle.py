"""CPU restatement (test infrastructure only) of the k-means the reference calls at
mmseg/models/decode_heads/cffm_head.py:280-282:

    kmeans = KMeans(n_clusters=self.n_clusters, max_iter=10, mode='euclidean', verbose=0)
    labels = kmeans.fit_predict(_c_cluster[ii]);  centers.append(kmeans.centroids)

``fast_pytorch_kmeans`` is a third-party dependency that is NOT vendored under /root/reference and is not
version-pinned anywhere in it (README.md:34 and requirements/ do not list it), and it is not installable here
(no network).  This file restates its published algorithm (fast_pytorch_kmeans/kmeans.py, ``fit_predict`` with
``mode='euclidean'``, ``minibatch=None``): similarity 2ab - |a|^2 - |b|^2, arg max, one-hot-mask means with NaN -> 0,
error = sum (c_new - c)^2, stop at error <= tol, labels of the last assignment.  The random initial centroids
(np.random.choice without replacement) make the library itself non-reproducible, and the reference's tests hold no
vector for it:  **parity unpinned**  -- the GPU path is compared with this restatement from identical initial centroids.
"""
import numpy as np
import torch


def euc_sim(a, b):
    return 2 * a @ b.transpose(-2, -1) - (a ** 2).sum(dim=1)[..., :, None] - (b ** 2).sum(dim=1)[..., None, :]


def fit_predict(X, n_clusters, max_iter=100, tol=1e-4, centroids=None):
    """X (N,E) fp32 CPU tensor -> (labels int64 (N,), centroids fp32 (K,E), iterations run)."""
    X = X.float()
    N = X.shape[0]
    if centroids is None:
        centroids = X[np.random.choice(N, size=[n_clusters], replace=False)]
    centroids = centroids.float().clone()
    closest, it = None, 0
    for it in range(1, max_iter + 1):
        closest = euc_sim(X, centroids).max(dim=-1)[1]
        mask = (closest[None].expand(n_clusters, -1) == torch.arange(n_clusters)[:, None]).float()
        c_grad = mask @ X / mask.sum(-1)[..., :, None]
        c_grad[c_grad != c_grad] = 0                                      # empty cluster: NaN -> 0
        error = (c_grad - centroids).pow(2).sum()
        centroids = c_grad                                                # lr = 1 without minibatches
        if error <= tol:
            break
    return closest, centroids, it


def inertia(X, centroids, labels):
    return ((X.float() - centroids.float()[labels]) ** 2).sum().item()
