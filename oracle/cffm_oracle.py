"""CPU oracle for the CFFM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch-on-CPU restatement of the reference algorithm, written from the
reference's behaviour (not copied): every function cites the reference file:line it
follows (paths relative to the reference root).  It is pinned against golden vectors
produced by executing the UNMODIFIED reference in the build container
(``oracle/make_goldens.py`` -> ``tests/golden/``; checked by ``tests/test_oracle_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module -- always as the checker or the CPU
baseline, never as the thing shipped.  The product package ``vss_cffm_b200`` never
imports it and has no CPU fallback.

The K/V assembling is restated as an explicit *coordinate gather* 
instead of the reference's roll / window_partition / unfold / cat pipeline, so the same
integer source-coordinate table is the spec of the CUDA gather kernel.

All functions take ``sd``: a mapping ``state-dict key -> torch.Tensor (fp32, CPU)`` with
the reference's key names, plus a key prefix.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- constants
# mmseg/models/backbones/mix_transformer.py:373-423
MIT_VARIANTS = {
    "mit_b0": dict(embed_dims=[32, 64, 160, 256], depths=[2, 2, 2, 2]),
    "mit_b1": dict(embed_dims=[64, 128, 320, 512], depths=[2, 2, 2, 2]),
    "mit_b2": dict(embed_dims=[64, 128, 320, 512], depths=[3, 4, 6, 3]),
    "mit_b3": dict(embed_dims=[64, 128, 320, 512], depths=[3, 4, 18, 3]),
    "mit_b4": dict(embed_dims=[64, 128, 320, 512], depths=[3, 8, 27, 3]),
    "mit_b5": dict(embed_dims=[64, 128, 320, 512], depths=[3, 6, 40, 3]),
}
MIT_HEADS = [1, 2, 5, 8]
MIT_SR = [8, 4, 2, 1]
# mmseg/models/decode_heads/cffm_head.py:74-95 (hard-coded hyper-parameters of the head)
CFFM_HEADS = 8
CFFM_WS = 7
CFFM_EXPAND = 3
CFFM_FOCAL_WINDOW = 5
CFFM_L_CLIPS = (1, 2, 3)
CFFM_K_CLIPS = (7, 5, 3)


def _ln(x, sd, p, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def resize(x, size):
    """mmseg/ops/wrappers.py:8-29 with mode='bilinear', align_corners=False."""
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)


# ============================================================================ MiT backbone
def mit_attention(sd, p, x, H, W, heads, sr):
    """Efficient (spatial-reduction) self-attention. mix_transformer.py:96-117."""
    B, N, C = x.shape
    d = C // heads
    q = _lin(x, sd, p + ".q").view(B, N, heads, d).transpose(1, 2)
    if sr > 1:
        xs = x.transpose(1, 2).reshape(B, C, H, W)
        xs = F.conv2d(xs, sd[p + ".sr.weight"], sd[p + ".sr.bias"], stride=sr)
        xs = xs.flatten(2).transpose(1, 2)
        xs = _ln(xs, sd, p + ".norm", 1e-5)          # bare nn.LayerNorm(dim): eps 1e-5 (:77)
    else:
        xs = x
    kv = _lin(xs, sd, p + ".kv").view(B, -1, 2, heads, d)
    k = kv[:, :, 0].transpose(1, 2)
    v = kv[:, :, 1].transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-2, -1)) * (d ** -0.5), dim=-1)   # scale after matmul (:109)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(o, sd, p + ".proj")


def mit_mixffn(sd, p, x, H, W):
    """fc1 -> depthwise 3x3 -> GELU(erf) -> fc2. mix_transformer.py:48-55,358-369."""
    B, N, _ = x.shape
    h = _lin(x, sd, p + ".fc1")
    Ch = h.shape[-1]
    h = h.transpose(1, 2).reshape(B, Ch, H, W)
    h = F.conv2d(h, sd[p + ".dwconv.dwconv.weight"], sd[p + ".dwconv.dwconv.bias"], padding=1, groups=Ch)
    h = F.gelu(h.flatten(2).transpose(1, 2))
    return _lin(h, sd, p + ".fc2")


def mit_forward(sd, prefix, img, variant):
    """MixVisionTransformer.forward_features, eval mode. mix_transformer.py:313-349.

    img (N,3,H,W) -> list of 4 NCHW feature maps at strides 4/8/16/32."""
    cfg = MIT_VARIANTS[variant]
    x = img
    outs = []
    for s in range(4):
        pe = f"{prefix}patch_embed{s + 1}"
        k, st = (7, 4) if s == 0 else (3, 2)
        x = F.conv2d(x, sd[pe + ".proj.weight"], sd[pe + ".proj.bias"], stride=st, padding=k // 2)
        Bn, C, H, W = x.shape
        x = _ln(x.flatten(2).transpose(1, 2), sd, pe + ".norm", 1e-5)    # bare LayerNorm (:175)
        for i in range(cfg["depths"][s]):
            bp = f"{prefix}block{s + 1}.{i}"
            x = x + mit_attention(sd, bp + ".attn", _ln(x, sd, bp + ".norm1", 1e-6), H, W, MIT_HEADS[s], MIT_SR[s])
            x = x + mit_mixffn(sd, bp + ".mlp", _ln(x, sd, bp + ".norm2", 1e-6), H, W)
        x = _ln(x, sd, f"{prefix}norm{s + 1}", 1e-6)
        x = x.reshape(Bn, H, W, C).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


# ====================================================================== CFFM index tables
def relative_position_index(q_win, k_win):
    """Closed form of get_relative_position_index. cffm_transformer.py:158-185.

    idx[q, k] = (qy - ky + kh - 1) * (qw + kw - 1) + (qx - kx + kw - 1), int64 (qh*qw, kh*kw)."""
    qh, qw = q_win
    kh, kw = k_win
    qy = torch.arange(qh).repeat_interleave(qw)
    qx = torch.arange(qw).repeat(qh)
    ky = torch.arange(kh).repeat_interleave(kw)
    kx = torch.arange(kw).repeat(kh)
    return (qy[:, None] - ky[None, :] + kh - 1) * (qw + kw - 1) + (qx[:, None] - kx[None, :] + kw - 1)


def ring_offsets(ws=CFFM_WS, e=CFFM_EXPAND):
    """(dr, dc) offsets, relative to the window origin, of the 4*ws*ws - 4*(ws-e)^2 'rolled'
    neighbour keys, in the order produced by ``valid_ind_rolled`` (cffm_transformer.py:280-285)
    applied to cat(tl, tr, bl, br) of the four torch.roll'ed maps (:389-416)."""
    offs = []
    for sy, sx in ((+e, +e), (+e, -e), (-e, +e), (-e, -e)):          # tl, tr, bl, br
        for r in range(ws):
            for c in range(ws):
                row_out = (r >= ws - e) if sy > 0 else (r < e)
                col_out = (c >= ws - e) if sx > 0 else (c < e)
                if row_out or col_out:
                    offs.append((r + sy, c + sx))
    return offs


def key_source_table(Hp, Wp, ws=CFFM_WS, e=CFFM_EXPAND, fw=CFFM_FOCAL_WINDOW,
                     l_clips=CFFM_L_CLIPS, k_clips=CFFM_K_CLIPS):
    """Integer source table of the assembled K/V sequence of every window.

    Returns (level, y, x) int64 tensors of shape (nW, N_keys).  level: 0 = full-resolution
    target map (Hp x Wp, cyclic), 1 = fc-pooled target windows, 2+k = fc-pooled reference
    frame k.  y = x = -1 where nn.Unfold zero-fills (outside the pooled grid), which is also
    where the -100 mask applies.  Key order = column order of the reference's logits
    (cffm_transformer.py:417-418,521): own window | ring | pooled target | ref 0 | ref 1 | ref 2.
    """
    nWh, nWw = Hp // ws, Wp // ws
    lev, ys, xs = [], [], []
    ring = ring_offsets(ws, e)
    for i in range(nWh):
        for j in range(nWw):
            L, Y, X = [], [], []
            for r in range(ws):                                           # own window (:378-383)
                for c in range(ws):
                    L.append(0); Y.append(ws * i + r); X.append(ws * j + c)
            for dr, dc in ring:                                           # cyclic ring (:389-416)
                L.append(0); Y.append((ws * i + dr) % Hp); X.append((ws * j + dc) % Wp)
            segs = [(1, 1, fw, nWh, nWw)]                                 # unfolds[0]: k=fw,s=1,p=fw//2 (:294-301)
            for k, (l, kc) in enumerate(zip(l_clips, k_clips)):           # unfolds_clips[k] (:333-343)
                segs.append((2 + k, l, kc, nWh * l, nWw * l))
            for level, stride, kc, gh, gw in segs:
                for u in range(kc):
                    for v in range(kc):
                        y = stride * i + u - kc // 2
                        x = stride * j + v - kc // 2
                        ok = 0 <= y < gh and 0 <= x < gw
                        L.append(level); Y.append(y if ok else -1); X.append(x if ok else -1)
            lev.append(L); ys.append(Y); xs.append(X)
    return torch.tensor(lev), torch.tensor(ys), torch.tensor(xs)


def cfm_bias_table(sd, p, ws=CFFM_WS, fw=CFFM_FOCAL_WINDOW, k_clips=CFFM_K_CLIPS):
    """Window-independent additive term of the logits, (nH, ws*ws, N_keys) fp32.
    cffm_transformer.py:536-587 (relative-position tables gathered by the index buffers)."""
    wa = ws * ws
    t0 = sd[p + ".relative_position_bias_table"]                           # (169, nH)
    parts = [t0[relative_position_index((ws, ws), (ws, ws)).reshape(-1)].view(wa, wa, -1).permute(2, 0, 1)]
    parts.append(sd[p + ".relative_position_bias_table_to_neighbors"][0])  # (nH, 49, 132)
    tw = sd[p + ".relative_position_bias_table_to_windows.0"]              # (nH, 121)
    parts.append(tw[:, relative_position_index((ws, ws), (fw, fw)).reshape(-1)].view(-1, wa, fw * fw))
    for k, kc in enumerate(k_clips):
        tk = sd[p + f".relative_position_bias_table_to_windows_clips.{k}"]
        parts.append(tk[:, relative_position_index((ws, ws), (kc, kc)).reshape(-1)].view(-1, wa, kc * kc))
    return torch.cat(parts, dim=2)


# =========================================================================== CFFA pooling
def fc_pool(x, w, b, k):
    """(B,H,W,C) -> (B,H/k,W/k,C): Linear(k*k -> 1) over each k x k patch, same weights for all
    channels. cffm_transformer.py:768-773 / 797-802 (window_partition_noreshape + pool layer)."""
    B, H, W, C = x.shape
    xp = x.view(B, H // k, k, W // k, k, C).permute(0, 1, 3, 5, 2, 4).reshape(B, H // k, W // k, C, k * k)
    return xp @ w.view(-1) + b.view(())


def cffa_assemble(sd, p, xn_pad, ws=CFFM_WS, l_clips=CFFM_L_CLIPS):
    """Coarse-to-fine feature assembling. cffm_transformer.py:739-805.

    xn_pad: (B, T, Hp, Wp, C) LayerNorm'ed, zero-padded frames (target = last).
    Returns [target pooled (B,nWh,nWw,C), ref0 pooled, ref1 pooled, ref2 pooled]."""
    B, T, Hp, Wp, C = xn_pad.shape
    nWh, nWw = Hp // ws, Wp // ws
    out = [fc_pool(xn_pad[:, -1], sd[p + ".pool_layers.0.weight"], sd[p + ".pool_layers.0.bias"], ws)]
    for k, l in enumerate(l_clips):
        wg = ws // l                                                      # floor(7/l) = 7, 3, 2 (:785)
        Hpool, Wpool = nWh * l * wg, nWw * l * wg                          # 63, 54, 54 at 60x60 (:786-790)
        xk = xn_pad[:, k]
        if (Hpool, Wpool) != (Hp, Wp):                                     # bilinear on the PADDED map (:794-795)
            xk = resize(xk.permute(0, 3, 1, 2), (Hpool, Wpool)).permute(0, 2, 3, 1)
        out.append(fc_pool(xk.contiguous(), sd[p + f".pool_layers_clips.{k}.weight"],
                           sd[p + f".pool_layers_clips.{k}.bias"], wg))
    return out


# ========================================================================== CFM attention
def cfm_attention(sd, p, xt_pad, pooled, heads=CFFM_HEADS, ws=CFFM_WS, return_probs=False):
    """Cross-frame feature mining attention, gather formulation. cffm_transformer.py:364-606.

    xt_pad (B,Hp,Wp,C): LN'ed zero-padded target map; pooled: output of cffa_assemble.
    Returns (B*nW, ws*ws, C) like the reference (window-major tokens)."""
    B, Hp, Wp, C = xt_pad.shape
    d = C // heads
    nWh, nWw = Hp // ws, Wp // ws
    nW, wa = nWh * nWw, ws * ws
    qkv = _lin(xt_pad, sd, p + ".qkv")                                     # (:374) pad rows give q=k=v=bias
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    # key bank per clip: [full-res target | pooled target | ref0 | ref1 | ref2 | one zero row]
    kb, vb, base = [k.reshape(B, -1, C)], [v.reshape(B, -1, C)], [0]
    widths = [Wp]
    for pm in pooled:                                                      # (:449-450, :495-496) K,V thirds only
        kv = _lin(pm, sd, p + ".qkv")
        base.append(base[-1] + kb[-1].shape[1])
        kb.append(kv[..., C:2 * C].reshape(B, -1, C))
        vb.append(kv[..., 2 * C:].reshape(B, -1, C))
        widths.append(pm.shape[2])
    zero_row = base[-1] + kb[-1].shape[1]
    kb = torch.cat(kb + [torch.zeros(B, 1, C)], dim=1)
    vb = torch.cat(vb + [torch.zeros(B, 1, C)], dim=1)
    lev, ys, xs = key_source_table(Hp, Wp, ws)
    flat = torch.tensor(base)[lev] + ys * torch.tensor(widths)[lev] + xs
    flat = torch.where(ys < 0, torch.full_like(flat, zero_row), flat)      # (nW, N)
    mask = torch.where(ys < 0, -100.0, 0.0)                                # (:445,:490) -100, not -inf
    N = flat.shape[1]
    k_all = kb[:, flat.reshape(-1)].view(B, nW, N, heads, d).permute(0, 1, 3, 2, 4)   # (B,nW,nH,N,d)
    v_all = vb[:, flat.reshape(-1)].view(B, nW, N, heads, d).permute(0, 1, 3, 2, 4)
    qw = q.view(B, nWh, ws, nWw, ws, heads, d).permute(0, 1, 3, 5, 2, 4, 6).reshape(B, nW, heads, wa, d)
    logits = (qw * (d ** -0.5)) @ k_all.transpose(-2, -1)                  # q*scale BEFORE matmul (:528-530)
    logits = logits + cfm_bias_table(sd, p, ws)[None, None] + mask[None, :, None, None, :]
    probs = torch.softmax(logits, dim=-1)
    o = (probs @ v_all).permute(0, 1, 3, 2, 4).reshape(B * nW, wa, C)      # (:601)
    o = _lin(o, sd, p + ".proj")
    if return_probs:
        return o, probs
    return o


def cffm_block(sd, p, x, heads=CFFM_HEADS, ws=CFFM_WS):
    """CffmTransformerBlock3d3.forward (shift 0, eval). cffm_transformer.py:709-832.
    x: (B,T,H,W,C) -> same shape; only the last (target) frame changes."""
    B, T, H, W, C = x.shape
    xn = _ln(x, sd, p + ".norm1", 1e-5)
    pb, pr = (ws - H % ws) % ws, (ws - W % ws) % ws
    xn = F.pad(xn, (0, 0, 0, pr, 0, pb))                                   # pad AFTER the norm (:716-724)
    Hp, Wp = H + pb, W + pr
    pooled = cffa_assemble(sd, p, xn, ws)
    aw = cfm_attention(sd, p + ".attn", xn[:, -1], pooled, heads, ws)      # (B*nW, 49, C)
    nWh, nWw = Hp // ws, Wp // ws
    a = aw.view(B, nWh, nWw, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)[:, :H, :W]  # (:812-821)
    t = x[:, -1] + a
    h = _ln(t, sd, p + ".norm2", 1e-5)
    t = t + _lin(F.gelu(_lin(h, sd, p + ".mlp.fc1")), sd, p + ".mlp.fc2")
    return torch.cat([x[:, :-1], t.unsqueeze(1)], dim=1)


def basic_layer3d3(sd, p, x, depth):
    """BasicLayer3d3.forward. cffm_transformer.py:917-927. x: (B,T,C,H,W) -> same."""
    x = x.permute(0, 1, 3, 4, 2)
    for i in range(depth):
        x = cffm_block(sd, f"{p}.blocks.{i}", x)
    return x.permute(0, 1, 4, 2, 3)


# ================================================================================== head
def _bn_eval(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=eps)


def head_mlp_decoder(sd, p, feats):
    """Per-frame SegFormer all-MLP decoder -> _c (N,256,h,w). cffm_head.py:102-119."""
    c1 = feats[0]
    n, _, h, w = c1.shape
    ups = []
    for i in (4, 3, 2, 1):
        c = feats[i - 1]
        t = _lin(c.flatten(2).transpose(1, 2), sd, f"{p}linear_c{i}.proj")
        t = t.permute(0, 2, 1).reshape(n, -1, c.shape[2], c.shape[3])
        ups.append(t if i == 1 else resize(t, (h, w)))
    y = F.conv2d(torch.cat(ups, dim=1), sd[p + "linear_fuse.conv.weight"])   # ConvModule: no conv bias with norm
    return F.relu(_bn_eval(y, sd, p + "linear_fuse.bn"))


def cffm_head_forward(sd, p, feats, batch_size, num_clips, cfg_num_clips, depth, return_intermediates=False):
    """CFFMHead_clips_resize1_8.forward in eval mode. cffm_head.py:99-157."""
    _c = head_mlp_decoder(sd, p, feats)
    n, C, h, w = _c.shape
    if num_clips != cfg_num_clips:                                          # early return (:127-129)
        x = F.conv2d(_c, sd[p + "linear_pred.weight"], sd[p + "linear_pred.bias"])
        return x.reshape(batch_size, num_clips, -1, h, w)[:, -1]
    h2, w2 = int(h / 2), int(w / 2)
    cf = resize(_c, (h2, w2)).reshape(batch_size, num_clips, C, h2, w2)      # (:131-135)
    c2 = basic_layer3d3(sd, p + "decoder_focal", cf, depth)                  # (:138)
    cat = torch.cat([cf[:, -1], c2[:, -1]], dim=1)
    x2 = F.conv2d(cat, sd[p + "linear_pred2.weight"], sd[p + "linear_pred2.bias"])
    x2 = resize(x2, (h, w))                                                  # (:145-149)
    if return_intermediates:
        return x2, dict(_c=_c, c_further=cf, c2=c2)
    return x2


def gene_prototype_head_forward(sd, p, feats, batch_size, num_clips, n_clusters=100, max_iter=10, init_centroids=None):
    """CFFMHead_clips_resize1_8_gene_prototype.forward (cffm_head.py:239-300): linear_pred logits of the last frame,
    the 1/8-scale clustering features (B, num_clips*h2*w2, C) and the k-means centres (B, K, C).  The k-means itself is
    the restatement in oracle/kmeans_oracle.py (third-party fast_pytorch_kmeans, parity unpinned)."""
    from . import kmeans_oracle
    _c = head_mlp_decoder(sd, p, feats)
    n, C, h, w = _c.shape
    x = F.conv2d(_c, sd[p + "linear_pred.weight"], sd[p + "linear_pred.bias"]).reshape(batch_size, num_clips, -1, h, w)
    assert batch_size == 1                                                   # (:269)
    h2, w2 = int(h / 2), int(w / 2)
    c2 = resize(_c, (h2, w2)).reshape(batch_size, num_clips, C, h2, w2)
    cl = c2.permute(0, 1, 3, 4, 2).reshape(batch_size, num_clips * h2 * w2, C)   # (:273-275)
    centers = []
    for ii in range(batch_size):
        init = None if init_centroids is None else init_centroids[ii]
        centers.append(kmeans_oracle.fit_predict(cl[ii], n_clusters, max_iter=max_iter, centroids=init)[1])
    return x[:, -1], cl, torch.stack(centers, dim=0)


# =============================================================================== CFFM++
def cluster_attention(sd, p, x, centers, heads=CFFM_HEADS):
    """WindowAttention_cluster.forward with only_use_cluster_center_as_context=True.
    swin_transformer_2d.py:208-262.  Per-token op: the window partition around it is a
    permutation that the reverse undoes, no positional bias or mask is applied."""
    B, N, C = x.shape
    d = C // heads
    q = F.linear(x, sd[p + ".qkv.weight"][:C], sd[p + ".qkv.bias"][:C]) * (d ** -0.5)
    q = q.view(B, N, heads, d).transpose(1, 2)
    kv = _lin(centers, sd, p + ".qkv_cluster").view(B, -1, 2, heads, d)
    k = kv[:, :, 0].transpose(1, 2)
    v = kv[:, :, 1].transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return _lin(o, sd, p + ".proj_cluster")


def cluster_layer(sd, p, x, centers, depth=1):
    """BasicLayer_cluster / SwinTransformerBlock_cluster, shift 0 for block 0.
    swin_transformer_2d.py:605-665,1103-1148.  x (B,H*W,C), centers (B,K,C)."""
    for i in range(depth):
        bp = f"{p}.blocks.{i}"
        xn = _ln(x, sd, bp + ".norm1", 1e-5)
        cn = _ln(centers, sd, bp + ".norm1", 1e-5)                           # same norm on prototypes (:622)
        x = x + cluster_attention(sd, bp + ".attn", xn, cn)
        x = x + _lin(F.gelu(_lin(_ln(x, sd, bp + ".norm2", 1e-5), sd, bp + ".mlp.fc1")), sd, bp + ".mlp.fc2")
    return x


def cffmpp_head_forward(sd, p, feats, centers, batch_size, num_clips, cfg_num_clips, depth):
    """CFFMHead_clips_resize1_8_finetune_w_prototype3.forward, eval, prototypes given as a
    tensor (the reference reads them from ./cluster_centers/<video>/centers.pt). cffm_head.py:423-535."""
    _c = head_mlp_decoder(sd, p, feats)
    n, C, h, w = _c.shape
    if num_clips != cfg_num_clips:
        x = F.conv2d(_c, sd[p + "linear_pred.weight"], sd[p + "linear_pred.bias"])
        return x.reshape(batch_size, num_clips, -1, h, w)[:, -1]
    h2, w2 = int(h / 2), int(w / 2)
    cf = resize(_c, (h2, w2)).reshape(batch_size, num_clips, C, h2, w2)
    c2 = basic_layer3d3(sd, p + "decoder_focal", cf, depth)
    x2 = F.conv2d(torch.cat([cf[:, -1], c2[:, -1]], dim=1), sd[p + "linear_pred2.weight"], sd[p + "linear_pred2.bias"])
    x2 = resize(x2, (h, w))
    tok = cf[:, -1].permute(0, 2, 3, 1).reshape(batch_size, h2 * w2, C)      # (:519)
    c3 = cluster_layer(sd, p + "decoder_swin", tok, centers, 1)
    c3 = c3.reshape(batch_size, h2, w2, C).permute(0, 3, 1, 2)
    x3 = F.conv2d(c3, sd[p + "linear_pred3.weight"], sd[p + "linear_pred3.bias"])
    x3 = resize(x3, (h, w))
    return x2 + 0.5 * x3                                                     # (:532)


# ============================================================================= segmentor
def segmentor_simple_test(sd, imgs, variant, depth, cfg_num_clips=4, centers=None, return_logits=False):
    """EncoderDecoder_clips.simple_test -> inference -> whole_inference -> encode_decode.
    mmseg/models/segmentors/encoder_decoder.py:367-378,502-572.  imgs: list of T (B,3,H,W)."""
    img = torch.stack(imgs, dim=1)
    B, T, _, H, W = img.shape
    feats = mit_forward(sd, "backbone.", img.reshape(B * T, 3, H, W), variant)
    if centers is None:
        logit = cffm_head_forward(sd, "decode_head.", feats, B, T, cfg_num_clips, depth)
    else:
        logit = cffmpp_head_forward(sd, "decode_head.", feats, centers, B, T, cfg_num_clips, depth)
    seg = resize(logit, (H, W))                                              # (:373-377)
    prob = torch.softmax(seg, dim=1)                                         # (:542)
    pred = prob.argmax(dim=1)                                                # (:564)
    if return_logits:
        return pred, logit
    return pred
