"""Integer tables of the CFFM window attention (host side, init time only).

Closed forms of the buffers the reference registers in ``WindowAttention3d3.__init__``
(mmseg/models/decode_heads/cffm_module/cffm_transformer.py:158-185, :267, :280-285, :316, :353);
tests/test_config_registry.py and tests/test_oracle_golden.py check them bit-exactly against the reference's own buffers
(tests/golden/index_tables.npz)."""
import torch

WS, EXPAND, FOCAL_WINDOW = 7, 3, 5          # hard-coded by the head (cffm_head.py:74-95)
L_CLIPS, K_CLIPS = (1, 2, 3), (7, 5, 3)
N_KEYS, N_RING = 289, 132


def relative_position_index(q_win, k_win):
    """idx[q,k] = (qy-ky+kh-1)*(qw+kw-1) + (qx-kx+kw-1): get_relative_position_index (:158-185)."""
    (qh, qw), (kh, kw) = q_win, k_win
    qy, qx = torch.arange(qh).repeat_interleave(qw), torch.arange(qw).repeat(qh)
    ky, kx = torch.arange(kh).repeat_interleave(kw), torch.arange(kw).repeat(kh)
    return (qy[:, None] - ky[None] + kh - 1) * (qw + kw - 1) + (qx[:, None] - kx[None] + kw - 1)


def valid_ind_rolled(ws=WS, e=EXPAND):
    """Indices, into cat(tl, tr, bl, br) of the four rolled 7x7 windows, of the tokens that fall
    outside the own window: the L-shaped strips (:280-285)."""
    idx = []
    for q, (down, right) in enumerate(((True, True), (True, False), (False, True), (False, False))):
        for r in range(ws):
            for c in range(ws):
                row_out = r >= ws - e if down else r < e
                col_out = c >= ws - e if right else c < e
                if row_out or col_out:
                    idx.append(q * ws * ws + r * ws + c)
    return torch.tensor(idx, dtype=torch.int64)


def ring_offsets(ws=WS, e=EXPAND):
    """(dy, dx), relative to the window origin, of the 132 ring keys in ``valid_ind_rolled`` order: window tl / tr /
    bl / br is the map rolled by (-e,-e) / (-e,+e) / (+e,-e) / (+e,+e) (cffm_transformer.py:389-400), so its token
    (r, c) sits at (r +- e, c +- e).  120 distinct offsets; 12 occur twice."""
    out = []
    for i in valid_ind_rolled(ws, e).tolist():
        q, r, c = i // (ws * ws), (i // ws) % ws, i % ws
        out.append((r + (e if q < 2 else -e), c + (e if q % 2 == 0 else -e)))
    return out


def assemble_bias_tc(table, to_neighbors, to_windows0, to_windows_clips, scale, layout, heads=8):
    """Additive logit term of the CFM kernel (``cffm_cfm_attention``): fp16 [heads, 49, pitch], DIVIDED by the attention
    scale, columns in the kernel's key order ``layout`` (= ``ops.cfm_layout()``): the 13 x 13 halo of the window
    row-major (own window from ``relative_position_bias_table``; ring from the dense neighbour table, the two entries
    of a ring key the reference lists twice folded into logaddexp(b1, b2), which leaves the softmax unchanged), then
    the four pooled windows (cffm_transformer.py:536-587).  Unused columns are -inf (their keys never count)."""
    wa, halo = WS * WS, WS + 2 * EXPAND
    dev = table.device
    out = torch.full((heads, wa, layout["pitch"]), float("-inf"), dtype=torch.float32, device=dev)
    h = torch.full((heads, wa, halo * halo), float("-inf"), dtype=torch.float32, device=dev)
    idx = relative_position_index((WS, WS), (WS, WS)).to(dev)
    own = table[idx.reshape(-1)].view(wa, wa, heads).permute(2, 0, 1)                     # [heads, q, own key n]
    for n in range(wa):
        h[:, :, (n // WS + EXPAND) * halo + n % WS + EXPAND] = own[:, :, n]
    nb = to_neighbors.reshape(heads, wa, N_RING).float()
    for m, (dy, dx) in enumerate(ring_offsets()):
        pos = (dy + EXPAND) * halo + dx + EXPAND
        h[:, :, pos] = torch.logaddexp(h[:, :, pos], nb[:, :, m])
    assert torch.isfinite(h).all()                                                        # every halo position is a key
    out[:, :, :halo * halo] = h
    idx = relative_position_index((WS, WS), (FOCAL_WINDOW, FOCAL_WINDOW)).to(dev)
    out[:, :, layout["rows"][1]:layout["rows"][1] + FOCAL_WINDOW ** 2] = to_windows0[:, idx.reshape(-1)].view(heads, wa, -1)
    for k, kc in enumerate(K_CLIPS):
        idx = relative_position_index((WS, WS), (kc, kc)).to(dev)
        r0 = layout["rows"][2 + k]
        out[:, :, r0:r0 + kc * kc] = to_windows_clips[k][:, idx.reshape(-1)].view(heads, wa, kc * kc)
    out = out / scale
    out[:, :, layout["npad"]:] = 0.0                                                      # beyond the key rows: never read
    return out.to(torch.float16).contiguous()
