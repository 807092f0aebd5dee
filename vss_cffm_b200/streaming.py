"""Streaming video inference with temporal re-use.

At test time the reference segments frame i of a video from the clip [i-9, i-6, i-3, i] (with the start-of-video
exceptions of mmseg/datasets/custom.py:2365-2386) and recomputes, for every target, the backbone, the MLP decoder and
the CFFA pooling of its three reference frames -- although each frame is a reference of up to three later targets and
a target itself.  Nothing in the CFFM blocks modifies a reference frame (cffm_transformer.py:826), so everything a
frame contributes as a reference is a pure function of that frame:

    per frame, once:  backbone -> folded MLP decoder -> 1/8-scale features x (cffm_head.py:102-133)
                      for every block and every temporal role k: LN(norm1) -> pad -> resize -> fc-pool -> K/V projection
                      (cffm_transformer.py:713-805, :495-518)          [2.3 MB of fp16 per frame for depth 2]
    per target:       the block loop on the target's own x with the cached K/V of its three references

``VideoStream.push(frame)`` does exactly that and returns the labels of the pushed frame; they are bit-identical to
``EncoderDecoder_clips`` run on the reference's clip for that frame (tests/test_streaming.py), at roughly a quarter of
the backbone work.  The reference's stateless API cannot express this; the stateless classes are untouched.
"""
import torch

from . import _abi, ops
from .workspace import Workspace

_H, _F = torch.float16, torch.float32
ROLE_TOKENS = (1, 4, 9)                                     # pooled tokens per window for reference role 0, 1, 2


def clip_indices(i, dilation=(-9, -6, -3)):
    """Frame indices of the clip the reference builds for frame i of a video (custom.py:2365-2386); the last one is i."""
    step = [i + d for d in dilation if i + d >= 0] + [i]
    if list(dilation) == [-9, -6, -3]:
        special = {3: [0, 1, 2, 3], 4: [0, 2, 3, 4], 5: [0, 2, 4, 5], 6: [0, 2, 4, 6], 7: [0, 3, 5, 7], 8: [0, 3, 6, 8]}
        step = special.get(i, step)
    return step


class VideoStream:
    """One stream per batch entry: ``push`` takes the next frame of each of the ``n_streams`` videos."""

    def __init__(self, model, n_streams=1, dilation=(-9, -6, -3), graph=False):
        """``graph=True``: once the history is full the per-frame step only depends on ``i mod history`` (which ring
        slots it reads and writes), so one CUDA graph per residue is captured and replayed; ``push`` then returns a
        static labels tensor that stays valid until the same residue comes round again (``history`` pushes later)."""
        head = model.decode_head
        if not hasattr(head, "decoder_focal") or getattr(head, "WITH_PROTOTYPES", False):
            raise _abi.CffmError("VideoStream drives the plain CFFM head (CFFMHead_clips_resize1_8)")
        self.model, self.head, self.B, self.dilation = model, head, n_streams, tuple(dilation)
        self.history = max(-min(dilation), 1) + 1               # frames a later target can still reference
        self.ws = Workspace()
        self.use_graph = graph
        self.ring = None                                         # [3 roles] of (history, depth, B, tokens, 2E) fp16
        self.reset()

    def reset(self):
        self.i = 0
        self.kv = {}                                             # frame index -> ring slot holding its reference K/V
        self.graphs = {}                                         # i mod history -> (CUDAGraph, static labels)
        self.static_in = None

    @torch.no_grad()
    def push(self, frames):
        """frames: (n_streams, 3, H, W) fp32 (CUDA, or host: copied) -> int64 labels (n_streams, H, W) of these frames."""
        i, hist = self.i, self.history
        steady = self.use_graph and i >= hist - 1 and list(self.dilation) == [-9, -6, -3]
        if not steady:
            labels = self._step(frames, i)
        else:
            dev = self.model._device()
            if self.static_in is None:
                self.static_in = torch.empty(tuple(frames.shape), dtype=_F, device=dev)
            self.static_in.copy_(frames, non_blocking=True)
            phase = i % hist
            if phase not in self.graphs:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    self._step(self.static_in, i)                # warm-up on a side stream (also fills this frame's ring slot)
                torch.cuda.current_stream(dev).wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = self._step(self.static_in, i)
                self.graphs[phase] = (g, out)
                for ws in [self.ws] + self.model.workspaces():   # the graphs replay on these addresses: never free them
                    ws.pin()
            g, labels = self.graphs[phase]
            g.replay()
        self.kv[i] = i % hist
        for old in [f for f in self.kv if f <= i - hist]:
            del self.kv[old]
        self.i += 1
        return labels

    def _step(self, frames, i):
        """All the device work of frame i (no host state is changed: capturable in a CUDA graph)."""
        model, head, ws, B = self.model, self.head, self.ws, self.B
        dev = model._device()
        frames = frames.to(dev, _F, non_blocking=True).contiguous()
        assert frames.dim() == 4 and frames.shape[0] == B
        H, W = frames.shape[-2:]
        P = head._plan or head._build_plan()
        E, HEADS = head.embed_dim, 8
        depth = len(P["blocks"])
        clip = clip_indices(i, self.dilation)
        hist = self.history
        early = len(clip) != head.num_clips                      # fewer than 3 references yet: eval early return (:127-129)
        # ---- this frame, once: backbone + folded MLP decoder
        feats = [head._as_nhwc16(t) for t in model.backbone(frames)]
        sizes = [(t.shape[1], t.shape[2]) for t in feats]
        h, w = sizes[0]
        proj = []
        for k, t in enumerate(feats):
            p = ws.get(f"p{k}", (B * sizes[k][0] * sizes[k][1], E), _H, device=dev)
            ops.gemm(t.reshape(-1, t.shape[3]), P["pw"][k], out16=p)
            proj.append(p)
        h2, w2 = h // 2, w // 2
        HW = h2 * w2
        Hp, Wp = (h2 + 6) // 7 * 7, (w2 + 6) // 7 * 7
        nW = (Hp // 7) * (Wp // 7)
        x32 = ws.get("x32", (B * HW, E), _F, device=dev)
        c16 = ws.get("c16", (B * HW, E), _H, device=dev)
        c_full = ws.get("c_full", (B * h * w, E), _H, device=dev) if early else None
        ops.head_fuse(proj, sizes, B, E, 0, P["shift"], c_full=c_full, half32=x32, half16=c16)
        # ---- what this frame contributes as a reference of later targets: K/V of every block for every role
        xn = ws.get("xn_ref", (B * HW, E), _H, device=dev)
        if self.ring is None or self.ring[0].shape[3] != nW:
            self.ring = [torch.zeros(hist, depth, B, ROLE_TOKENS[k] * nW, 2 * E, dtype=_H, device=dev) for k in range(3)]
        roles = [self.ring[k][i % hist] for k in range(3)]
        for bi, b in enumerate(P["blocks"]):
            ops.cffa_norm_frames(x32, b["n1g"], b["n1b"], b["n1eps"], xn, None, B, B, h2, w2, Hp, Wp, E)
            for k in range(3):
                n = ROLE_TOKENS[k] * nW
                pooled = ws.get(f"pooled_r{k}", (B * n, E), _H, device=dev)
                ops.cffa_pool_level(xn, B, k + 1, h2, w2, E, b["pool_w"], b["pool_b"], pooled)
                ops.gemm(pooled, b["qkv_w"][E:], bias=b["qkv_b"][E:], out16=roles[k][bi].view(B * n, 2 * E))
        labels = torch.empty(B, H, W, dtype=torch.int64, device=dev)
        ncls = model.num_classes
        if early:
            lg = ws.get("lg_full", (B * h * w, P["ncp"]), _F, device=dev)
            ops.gemm(c_full, P["pred_w"], bias=P["pred_b"], out32=lg)
            self._tail(lg, labels, B, h, w, h, w, H, W, ncls)
            return labels
        # ---- target: CFFM blocks with the cached reference K/V (cffm_transformer.py:709-832)
        xt, ct16 = x32, c16
        xn_t = ws.get("xn_t", (B * HW, E), _H, device=dev)
        xt_pad = ws.get("xt_pad", (ops.apron_rows(B, h2, w2), E), _H, device=dev)
        pooled_t = ws.get("pooled_t", (B * nW, E), _H, device=dev)
        kv_t = ws.get("kv_t", (B * nW, 2 * E), _H, device=dev)
        qkv_t = ws.get("qkv_t", (ops.apron_rows(B, h2, w2), 3 * E), _H, device=dev)
        kvp = ws.get("kvp", (B, 15 * nW, 2 * E), _H, device=dev)
        ao = ws.get("ao", (B * HW, E), _H, device=dev)
        xn2 = ws.get("xn2", (B * HW, E), _H, device=dev)
        hid = ws.get("hid", (B * HW, 4 * E), _H, device=dev)
        xt16 = ws.get("xt16", (B * HW, E), _H, device=dev)
        for bi, b in enumerate(P["blocks"]):
            ops.cffa_norm_frames(xt, b["n1g"], b["n1b"], b["n1eps"], xn_t, xt_pad, B, 0, h2, w2, Hp, Wp, E)
            ops.cffa_pool_level(xn_t, B, 0, h2, w2, E, b["pool_w"], b["pool_b"], pooled_t)
            ops.gemm(xt_pad, b["qkv_w"], bias=b["qkv_b"], out16=qkv_t)
            ops.gemm(pooled_t, b["qkv_w"][E:], bias=b["qkv_b"][E:], out16=kv_t)
            kvp[:, :nW].copy_(kv_t.view(B, nW, 2 * E))
            off = nW
            for k, f in enumerate(clip[:-1]):                    # role k = position in the clip
                n = ROLE_TOKENS[k] * nW
                kvp[:, off:off + n].copy_(self.ring[k][f % hist, bi])
                off += n
            ops.cfm_attention(qkv_t, kvp.view(B * 15 * nW, 2 * E), b["bias"], ao, B, h2, w2, E, HEADS, (E // HEADS) ** -0.5)
            ops.gemm(ao, b["proj_w"], bias=b["proj_b"], residual=xt, out32=xt)
            ops.layernorm(xt, b["n2g"], b["n2b"], b["n2eps"], out16=xn2)
            ops.gemm(xn2, b["f1w"], bias=b["f1b"], out16=hid, act=ops.ACT_GELU)
            ops.gemm(hid, b["f2w"], bias=b["f2b"], residual=xt, out32=xt, out16=xt16 if bi == depth - 1 else None)
        lg = ws.get("lg", (B * HW, P["ncp"]), _F, device=dev)
        ops.gemm(ct16, P["pred2_w"][:, :E], bias=P["pred2_b"], out32=lg)
        ops.gemm(xt16, P["pred2_w"][:, E:], residual=lg, out32=lg)
        self._tail(lg, labels, B, h2, w2, h, w, H, W, ncls)
        return labels

    @staticmethod
    def _tail(lg, labels, B, hs, ws_, h, w, H, W, ncls):
        """The head's resize to 1/4 scale + the segmentor's resize to the input size + arg max (same calls as
        EncoderDecoder_clips.labels_from_frames)."""
        if ops.upsample2_argmax_supported(h, w, H, W):
            ops.upsample2_argmax(lg, ncls, labels, B, hs, ws_, h, w, H, W)
        else:
            logits = torch.empty(B, ncls, h, w, dtype=_F, device=lg.device)
            ops.resize_nhwc_to_nchw(lg, ncls, logits, B, hs, ws_, h, w)
            ops.resize_argmax(logits, labels, B, ncls, h, w, H, W)
