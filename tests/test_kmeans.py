"""k-means prototypes of CFFM++ (reference call site cffm_head.py:267-294).

CPU: properties of the oracle restatement (the third-party library is un-pinned and absent: parity unpinned, no golden
vector exists).  GPU: the tensor-core path against that restatement from identical initial centroids, and the
prototype-generation head against the oracle head."""
import os

import numpy as np
import pytest
import torch

from oracle import cffm_oracle as O
from oracle import kmeans_oracle as KO
from vss_cffm_b200 import synth


def blobs(n_per, k, e, seed, spread=0.05):
    r = np.random.RandomState(seed)
    means = r.standard_normal((k, e)).astype(np.float32) * 2
    x = np.concatenate([means[j] + spread * r.standard_normal((n_per, e)).astype(np.float32) for j in range(k)])
    perm = r.permutation(len(x))
    lab = np.repeat(np.arange(k), n_per)[perm]
    return torch.from_numpy(x[perm]), torch.from_numpy(means), torch.from_numpy(lab)


# ------------------------------------------------------------------------------------------ oracle (CPU)
def test_oracle_recovers_separated_blobs():
    x, means, lab = blobs(50, 6, 32, 0)
    init = x[[int((lab == j).nonzero()[0]) for j in range(6)]]              # one seed point per blob
    labels, cen, it = KO.fit_predict(x, 6, max_iter=10, centroids=init)
    assert torch.equal(labels, lab)
    assert (cen - means).abs().max() < 0.05
    assert it <= 3                                                            # error <= tol stops the loop


def test_oracle_labels_belong_to_the_last_assignment_and_empty_clusters_become_zero():
    x, _, _ = blobs(40, 4, 16, 1)
    init = torch.cat([x[:4], torch.full((1, 16), 100.0)])                   # the 5th centre attracts nothing
    labels, cen, it = KO.fit_predict(x, 5, max_iter=1, centroids=init)
    assert it == 1 and torch.equal(labels, KO.euc_sim(x, init).max(dim=-1)[1])   # assignment precedes the update
    assert (labels != 4).all() and cen[4].abs().sum() == 0                   # NaN of the empty mean -> 0


def test_oracle_random_initialisation_draws_distinct_points():
    x, _, _ = blobs(30, 3, 8, 2)
    np.random.seed(5)
    idx = np.random.choice(90, size=[7], replace=False)
    np.random.seed(5)
    labels, cen, it = KO.fit_predict(x, 7, max_iter=0 + 1)
    assert len(set(idx.tolist())) == 7 and labels.shape == (90,) and cen.shape == (7, 8)


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def KMeans():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vss_cffm_b200.kmeans import KMeans as K
    torch.set_grad_enabled(False)
    return K


@pytest.mark.gpu
def test_gpu_kmeans_blobs_exact(KMeans):
    x, means, lab = blobs(200, 10, 256, 3)
    x = x.half().float()                                                      # the GPU path clusters the fp16 rounding
    init = x[[int((lab == j).nonzero()[0]) for j in range(10)]]
    ref_lab, ref_cen, ref_it = KO.fit_predict(x, 10, max_iter=10, centroids=init)
    km = KMeans(n_clusters=10, max_iter=10, mode="euclidean")
    got = km.fit_predict(x.cuda(), centroids=init.cuda())
    assert torch.equal(got.cpu(), ref_lab) and km.n_iter_ == ref_it
    assert (km.centroids.cpu() - ref_cen).abs().max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,E,iters", [(14400, 100, 256, 10), (3600, 100, 256, 10), (1001, 37, 64, 5)])
def test_gpu_kmeans_random_features_vs_oracle(KMeans, N, K, E, iters):
    """No cluster structure (the hard case: many near-ties).  Same initial centroids -> same partition quality; the few
    points that sit within fp32 rounding of a cell boundary may be assigned differently."""
    x = synth.synth_array((N, E), 77).half().float()
    np.random.seed(11)
    init = x[np.random.choice(N, size=[K], replace=False)]
    ref_lab, ref_cen, _ = KO.fit_predict(x, K, max_iter=iters, centroids=init)
    km = KMeans(n_clusters=K, max_iter=iters, mode="euclidean")
    got = km.fit_predict(x.cuda(), centroids=init.cuda()).cpu()
    agree = (got == ref_lab).float().mean().item()
    i_ref, i_got = KO.inertia(x, ref_cen, ref_lab), KO.inertia(x, km.centroids.cpu(), got)
    print(f"kmeans N={N} K={K}: label agreement {agree:.4f}, inertia {i_got:.6e} vs {i_ref:.6e}")
    assert agree >= 0.99, agree
    assert abs(i_got - i_ref) <= 1e-3 * i_ref
    # first iteration alone is a pure function of (x, init): compare it tightly
    l1, c1, _ = KO.fit_predict(x, K, max_iter=1, centroids=init)
    km1 = KMeans(n_clusters=K, max_iter=1)
    g1 = km1.fit_predict(x.cuda(), centroids=init.cuda()).cpu()
    assert (g1 == l1).float().mean().item() >= 0.9995
    same = torch.stack([(g1 == j).sum() == (l1 == j).sum() for j in range(K)])
    assert (km1.centroids.cpu()[same] - c1[same]).abs().max() < 1e-4


@pytest.mark.gpu
def test_gpu_kmeans_empty_cluster_and_random_init(KMeans):
    x, _, _ = blobs(100, 4, 64, 4)
    x = x.half().float()
    init = torch.cat([x[:4], torch.full((1, 64), 100.0)])
    km = KMeans(n_clusters=5, max_iter=3)
    lab = km.fit_predict(x.cuda(), centroids=init.cuda()).cpu()
    assert (lab != 4).all() and km.centroids[4].abs().sum().item() == 0
    np.random.seed(3)
    km2 = KMeans(n_clusters=8, max_iter=10)
    lab2 = km2.fit_predict(x.cuda())
    np.random.seed(3)
    ref_lab, ref_cen, _ = KO.fit_predict(x, 8, max_iter=10)                   # same numpy stream -> same initial points
    assert (lab2.cpu() == ref_lab).float().mean().item() >= 0.99
    with pytest.raises(Exception):
        KMeans(n_clusters=5, mode="cosine")
    with pytest.raises(Exception):
        km.fit_predict(x)                                                     # CPU tensor: no fallback


@pytest.mark.gpu
def test_gpu_gene_prototype_head_vs_oracle(KMeans, tmp_path, golden_dir):
    """CFFMHead_clips_resize1_8_gene_prototype through the registry: logits of the last frame and clustering features vs
    the oracle head; centres vs the oracle k-means run on the SAME features and initial points; centers.pt is written
    where the CFFM++ head reads it (cffm_head.py:286-294, :429-433)."""
    import vss_cffm_b200 as V
    m = V.build_segmentor(V.model_cfg("b0", "proto"))
    synth.fill_module(m, 12)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    head = m.decode_head
    head.save_path = str(tmp_path) + "/"
    B, T, H, W = 1, 4, 96, 160                                                # 4 x 12 x 20 = 960 points >= 100 clusters
    imgs = synth.synth_clip(B, T, H, W, seed=12)
    metas = [synth.img_metas(B, H, W, video="vidA")]
    np.random.seed(21)
    pred = m(img=[imgs], img_metas=metas, return_loss=False)
    assert len(pred) == 1 and pred[0].shape == (H, W)
    saved = torch.load(os.path.join(str(tmp_path), "vidA", "centers.pt"))
    assert tuple(saved.shape) == (1, 100, 256) and saved.dtype == torch.float32
    assert torch.equal(saved, head.centers.cpu())
    # oracle head on the oracle's own fp32 backbone features
    feats = O.mit_forward(sd, "backbone.", torch.stack(imgs, dim=1).reshape(B * T, 3, H, W), "mit_b0")
    ref_logit, ref_feat, _ = O.gene_prototype_head_forward(sd, "decode_head.", feats, B, T, n_clusters=100, max_iter=1,
                                                            init_centroids=torch.zeros(1, 100, 256))
    frames, _, _ = m._stack(imgs)
    np.random.seed(21)
    logits = m.encode_decode_frames(frames, metas[0], B, T, save=False).float().cpu()
    e = ((logits - ref_logit).abs().max() / ref_logit.abs().max()).item()
    x, t = m._features(frames, B, T)
    _, gfeat, _ = head.cluster_features(x, B, t, frame_major=True)
    gfeat = gfeat.float().cpu().permute(1, 0, 2, 3).reshape(1, -1, 256)
    ef = ((gfeat - ref_feat).abs().max() / ref_feat.abs().max()).item()
    print(f"gene_prototype: logits rel err {e:.2e}, clustering features rel err {ef:.2e}")
    assert e <= 1e-2 and ef <= 1e-2
    # k-means proper: same features, same numpy stream
    np.random.seed(21)
    ref_lab, ref_cen, _ = KO.fit_predict(gfeat[0], 100, max_iter=10)
    i_ref = KO.inertia(gfeat[0], ref_cen, ref_lab)
    got_lab = KO.euc_sim(gfeat[0], saved[0]).max(dim=-1)[1]
    i_got = KO.inertia(gfeat[0], saved[0], got_lab)
    assert abs(i_got - i_ref) <= 2e-3 * i_ref, (i_got, i_ref)
