"""Integer tables of the CFFM window attention (host side, init time only).

Closed forms of the buffers the reference registers in ``WindowAttention3d3.__init__``
(mmseg/models/decode_heads/cffm_module/cffm_transformer.py:158-185, :267, :280-285, :316, :353);
tests/test_tables.py checks them bit-exactly against the reference's own buffers
(tests/golden/index_tables.npz)."""
import torch

WS, EXPAND, FOCAL_WINDOW = 7, 3, 5          # hard-coded by the head (cffm_head.py:74-95)
L_CLIPS, K_CLIPS = (1, 2, 3), (7, 5, 3)
N_KEYS, N_KEYS_PAD, N_RING = 289, 320, 132


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


def assemble_bias(table, to_neighbors, to_windows0, to_windows_clips, heads=8):
    """Window-independent additive logit term, fp32 [heads, 64, 320] (rows >= 49 / cols >= 289 zero):
    the six in-place slice adds of cffm_transformer.py:536-587 gathered once at plan time."""
    wa = WS * WS
    out = torch.zeros(heads, 64, N_KEYS_PAD, dtype=torch.float32, device=table.device)
    dev = table.device
    idx = relative_position_index((WS, WS), (WS, WS)).to(dev)
    out[:, :wa, 0:49] = table[idx.reshape(-1)].view(wa, wa, heads).permute(2, 0, 1)
    out[:, :wa, 49:181] = to_neighbors.reshape(heads, wa, N_RING)
    idx = relative_position_index((WS, WS), (FOCAL_WINDOW, FOCAL_WINDOW)).to(dev)
    out[:, :wa, 181:206] = to_windows0[:, idx.reshape(-1)].view(heads, wa, FOCAL_WINDOW ** 2)
    col = 206
    for k, kc in enumerate(K_CLIPS):
        idx = relative_position_index((WS, WS), (kc, kc)).to(dev)
        out[:, :wa, col:col + kc * kc] = to_windows_clips[k][:, idx.reshape(-1)].view(heads, wa, kc * kc)
        col += kc * kc
    assert col == N_KEYS
    return out.contiguous()
