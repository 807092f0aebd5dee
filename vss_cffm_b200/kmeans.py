"""k-means for the CFFM++ prototypes, same call surface as the class the reference uses
(``fast_pytorch_kmeans.KMeans``, call site cffm_head.py:280-282:
``KMeans(n_clusters=100, max_iter=10, mode='euclidean', verbose=0).fit_predict(x)`` then ``.centroids``).

The library is an un-pinned third-party dependency that is not part of the reference tree; its published
algorithm (Lloyd iterations, ``mode='euclidean'``, no minibatch) is what is implemented: random initial
centroids drawn with ``np.random.choice(N, K, replace=False)``, then per iteration

    closest  = argmax_j (2 x.c_j - |x|^2 - |c_j|^2)
    c_new[j] = mean of the members of j   (0 for an empty cluster: the NaN of 0/0 is zeroed)
    error    = sum (c_new - c)^2 ;  c <- c_new ;  stop when error <= tol (1e-4)

and the labels returned are the ones of the LAST assignment (taken before the final update).

Both contractions run on the tensor cores (vss_cffm_b200/csrc/kmeans.cu explains the hi/lo split that keeps the
fp32 centroids at ~22 bits inside fp16 MMAs); no CPU fallback.
"""
import numpy as np
import torch

from . import _abi, ops

_H, _F = torch.float16, torch.float32


def _round_up(x, m):
    return (x + m - 1) // m * m


class KMeans:
    def __init__(self, n_clusters, max_iter=100, tol=0.0001, verbose=0, mode="euclidean", minibatch=None):
        if mode != "euclidean":
            raise _abi.CffmError("only mode='euclidean' (the one the reference uses, cffm_head.py:280) is implemented")
        if minibatch is not None:
            raise _abi.CffmError("minibatch k-means is not used by the reference and is not implemented")
        if n_clusters > 256:
            raise _abi.CffmError("at most 256 clusters (the reference uses 100, cffm_head.py:217)")
        self.n_clusters, self.max_iter, self.tol, self.verbose = n_clusters, max_iter, tol, verbose
        self.centroids = None
        self.n_iter_ = 0

    def fit_predict(self, X, centroids=None):
        """X: (N, E) CUDA tensor (fp16 or fp32; the tensor-core path works on its fp16 rounding).  Returns int64 labels (N,);
        ``self.centroids`` is the fp32 (K, E) result.  ``centroids``: optional (K, E) initial centres."""
        if not X.is_cuda:
            raise _abi.CffmError("KMeans.fit_predict: expected a CUDA tensor (no CPU fallback exists)")
        _abi.require_device()
        N, E = X.shape
        K = self.n_clusters
        if N < K:
            raise ValueError(f"{N} points for {K} clusters")
        if E % 8:
            raise _abi.CffmError(f"feature width must be a multiple of 8, got {E}")
        dev = X.device
        x16 = X.detach().to(_H).contiguous()
        if centroids is None:
            idx = np.random.choice(N, size=[K], replace=False)           # the library's initialisation
            cen = x16[torch.from_numpy(idx).to(dev)].to(_F).contiguous()
        else:
            cen = centroids.detach().to(dev, _F).contiguous().clone()
            assert tuple(cen.shape) == (K, E)
        Kp, Np = _round_up(K, 128), _round_up(N, 8)
        xt = torch.empty(E, Np, dtype=_H, device=dev)
        _abi.call("cffm_transpose_f16", x16.data_ptr(), N, E, xt.data_ptr(), Np, ops._stream())
        c_hi = torch.empty(Kp, E, dtype=_H, device=dev)
        c_lo = torch.empty(Kp, E, dtype=_H, device=dev)
        cnorm = torch.empty(Kp, dtype=_F, device=dev)
        scores = torch.empty(N, Kp, dtype=_F, device=dev)
        labels = torch.empty(N, dtype=torch.int64, device=dev)
        onehot = torch.empty(Kp, Np, dtype=_H, device=dev)
        counts = torch.zeros(K, dtype=torch.int32, device=dev)
        S = ops.splitk_plan(Kp, E, Np)
        partials = torch.empty(S, Kp, E, dtype=_F, device=dev)
        err_pc = torch.zeros(Kp, dtype=_F, device=dev)
        error = torch.zeros(1, dtype=_F, device=dev)
        done = torch.zeros(1, dtype=torch.int32, device=dev)
        _abi.call("cffm_kmeans_prepare", cen.data_ptr(), c_hi.data_ptr(), c_lo.data_ptr(), cnorm.data_ptr(), K, Kp, E, ops._stream())
        self.n_iter_ = 0
        for _ in range(self.max_iter):
            ops.gemm(x16, c_hi, out32=scores)
            ops.gemm(x16, c_lo, residual=scores, out32=scores)
            counts.zero_()
            _abi.call("cffm_kmeans_assign", scores.data_ptr(), Kp, cnorm.data_ptr(), N, Np, K, Kp, labels.data_ptr(),
                      onehot.data_ptr(), counts.data_ptr(), ops._stream())
            ops.gemm_splitk(onehot, xt, partials)
            _abi.call("cffm_kmeans_update", partials.data_ptr(), S, counts.data_ptr(), cen.data_ptr(), c_hi.data_ptr(),
                      c_lo.data_ptr(), cnorm.data_ptr(), err_pc.data_ptr(), error.data_ptr(), done.data_ptr(), K, Kp, E,
                      ops._stream())
            self.n_iter_ += 1
            if float(error.item()) <= self.tol:                         # the library's stop test (one host read per iteration)
                break
        self.centroids = cen
        return labels
