"""Multi-GPU execution of the CFFM hot path (one process per GPU, torch.distributed / NCCL over NVLink).

Two ways the path shards:

1. **Clips** are independent in eval mode (BatchNorm uses running statistics, nothing mixes clips): shard the clip
   batch across ranks, no data-path collective.  This is the reference's own strategy (DistributedSampler + DDP,
   tools/test.py:123-149) and what ``bench.py --gpus N`` measures by default (``shard_clips``).

2. **Frames**: backbone + MLP decoder are per frame; only the CFFM blocks mix the frames of a clip, and only by
   reading, for every block, the pooled K/V of the three reference frames (cffm_transformer.py:780-805, :470-518).
   Reference frames are never modified by a block (:826), so the owner of reference frame (b, t) can compute
   LN -> role-t pooling -> K/V projection for ALL blocks up front.  ``FrameShardedRunner`` does exactly that, then
   issues ONE all-gather of the packed K/V (per rank and block: its role-0 / 1 / 2 maps of 81 / 324 / 729 tokens x 512 fp16
   back to back, nothing padded) on a side stream, and the owners of the target frames run the CFM attention + FFN
   locally; the CFM kernel reads the gathered buffer in place through a per-clip (slot, rank) table.
   The payload is ~1.2 MB per clip and block: latency-bound on NVLink, so it is a single NCCL call hidden behind the
   target frames' norm1 / pooling / QKV GEMM, not a fused kernel.

Ownership: frame (b, t) of a global batch of Bg clips lives on rank (b + t) mod G when Bg >= G (every rank then owns
Bg/G frames of each temporal role, i.e. Bg/G targets: balanced) and on rank (b*T + t) mod G otherwise.
"""
import torch
import torch.distributed as dist

from . import _abi, ops
from .workspace import Workspace

_H, _F = torch.float16, torch.float32
ROLE_TOKENS = (1, 4, 9)                                     # pooled tokens per window for reference role 0, 1, 2


def shard_clips(n_clips, world, rank):
    """Contiguous clip range of ``rank`` (clip sharding, no collective)."""
    per, rem = divmod(n_clips, world)
    lo = rank * per + min(rank, rem)
    return range(lo, lo + per + (1 if rank < rem else 0))


class FrameShardPlan:
    """Static ownership map of the frames of a global batch: pure host logic (tested on CPU with gloo)."""

    def __init__(self, n_clips, n_frames, world):
        if n_frames != 4:
            raise _abi.CffmError("CFFM needs exactly 3 reference frames + 1 target per clip (cffm_head.py:93-94)")
        self.Bg, self.T, self.G = n_clips, n_frames, world
        self.frames = [[] for _ in range(world)]            # rank -> [(b, t)] frame-major (sorted by t, then b)
        for t in range(n_frames):
            for b in range(n_clips):
                self.frames[self.owner(b, t)].append((b, t))
        self.refs = [[f for f in fr if f[1] < n_frames - 1] for fr in self.frames]
        self.targets = [[b for b, t in fr if t == n_frames - 1] for fr in self.frames]
        self.max_slots = max(1, max(len(r) for r in self.refs))
        self._slot = {f: (rank, i) for rank, r in enumerate(self.refs) for i, f in enumerate(r)}
        # per temporal role k: how many reference frames a rank owns at most, and where frame (b, k) sits among its owner's
        # role-k frames.  Every rank packs its role-k K/V maps back to back (role_slots[k] maps of {1,4,9}[k] nW tokens), so the
        # all-gather moves exactly the tokens the CFM kernel reads (no padding of role 0 / 1 frames to the role-2 size).
        self.role_count = [[sum(1 for f in r if f[1] == k) for k in range(n_frames - 1)] for r in self.refs]
        self.role_slots = [max(1, max(c[k] for c in self.role_count)) for k in range(n_frames - 1)]
        self._role_index = {}
        for rank, r in enumerate(self.refs):
            seen = [0] * (n_frames - 1)
            for (b, t) in r:
                self._role_index[(b, t)] = (seen[t], rank)
                seen[t] += 1

    def owner(self, b, t):
        return (b + t) % self.G if self.Bg >= self.G else (b * self.T + t) % self.G

    def slot_of(self, b, t):
        """(owner rank, slot index in that rank's send buffer) of reference frame (b, t)."""
        return self._slot[(b, t)]

    def role_index(self, b, t):
        """(index among the owner's role-t frames, owner rank) of reference frame (b, t)."""
        return self._role_index[(b, t)]

    def role_offsets(self, nW):
        """Token offset of every role's block in a rank's packed buffer, and the buffer's token count."""
        off, o = [], 0
        for k, per in enumerate(ROLE_TOKENS):
            off.append(o)
            o += self.role_slots[k] * per * nW
        return off, o

    def gathered_index(self, b, t):
        """Row of reference frame (b, t) in the all-gathered buffer [G * max_slots, ...]."""
        rank, slot = self._slot[(b, t)]
        return rank * self.max_slots + slot


def all_gather_slots(send, group=None, out=None):
    """One all-gather of every rank's [max_slots, ...] buffer -> [G * max_slots, ...] (NCCL on GPUs, gloo in tests)."""
    world = dist.get_world_size(group)
    recv = out if out is not None else send.new_empty((world * send.shape[0],) + tuple(send.shape[1:]))
    if dist.get_backend(group) == "gloo":
        dist.all_gather(list(recv.chunk(world, dim=0)), send, group=group)
    else:
        dist.all_gather_into_tensor(recv, send, group=group)
    return recv


def assemble_kv(plan, gathered, clip, block, nW, out):
    """Host-side reference of what the slot table means (the runner itself copies nothing: ``cffm_cfm_attention_slots`` reads
    the gathered buffer in place).  Copy the K/V of clip ``clip``'s three reference frames for ``block`` from the gathered buffer
    [G*max_slots, depth, 9 nW, 2C] into ``out`` [15 nW, 2C] behind the nW pooled-target rows (layout of
    cffm_cfm_attention's kv_pooled: target | ref0 | ref1 | ref2)."""
    off = nW
    for t, per in enumerate(ROLE_TOKENS):
        n = per * nW
        out[off:off + n].copy_(gathered[plan.gathered_index(clip, t), block, :n])
        off += n
    return out


class FrameShardedRunner:
    """Runs EncoderDecoder_clips inference with the frames of a global clip batch sharded over the ranks."""

    def __init__(self, model, plan, rank, group=None):
        self.model, self.plan, self.rank, self.group = model, plan, rank, group
        self.ws = Workspace()

    def _slot_table(self, dev):
        """int32 [n_targets, 3, 2] on the device: (index among the owner's role-t maps, owner rank) of reference frame t of each
        target clip of this rank."""
        if getattr(self, "_slots", None) is None or self._slots.device != dev:
            rows = [[list(self.plan.role_index(clip, t)) for t in range(3)] for clip in self.plan.targets[self.rank]]
            self._slots = torch.tensor(rows, dtype=torch.int32).reshape(-1, 3, 2).to(dev)
        return self._slots

    def local_frames(self):
        """[(clip, t)] this rank must be fed, in the order ``run`` expects them."""
        return list(self.plan.frames[self.rank])

    def run(self, frames):
        """frames: (n_local, 3, H, W) fp32 on this rank's device, ordered as ``local_frames()``.
        Returns int64 labels (n_targets, H, W) for the clips ``plan.targets[rank]`` (may be empty)."""
        model, plan, ws = self.model, self.plan, self.ws
        head = model.decode_head
        P = head._plan or head._build_plan()
        E, HEADS = head.embed_dim, 8
        mine = plan.frames[self.rank]
        n_loc, n_ref, n_t = len(mine), len(plan.refs[self.rank]), len(plan.targets[self.rank])
        assert frames.shape[0] == n_loc
        H, W = frames.shape[-2:]
        dev = frames.device
        depth = len(P["blocks"])
        # ---- per-frame work: backbone + folded MLP decoder (cffm_head.py:102-133)
        if n_loc:
            proj = [None] * 4

            def project_stage(i, feat):
                """Folded linear_c{i+1} of ONE stage, on a side stream the moment that stage exists (beside the later stages)."""
                t = head._as_nhwc16(feat)
                proj[i] = ws.get(f"p{i}", (t.shape[0] * t.shape[1] * t.shape[2], E), _H, device=dev)
                with ops.fork("proj"):
                    ops.gemm(t.reshape(-1, t.shape[3]), P["pw"][i], out16=proj[i])

            feats = [head._as_nhwc16(t) for t in model.backbone(frames, stage_hook=project_stage)]
            sizes = [(t.shape[1], t.shape[2]) for t in feats]
            h, w = sizes[0]
            ops.join("proj")
        else:
            h, w = (H - 1) // 4 + 1, (W - 1) // 4 + 1              # OverlapPatchEmbed k7 s4 p3 (mix_transformer.py:173-195): same on every rank
        h2, w2 = h // 2, w // 2
        HW = h2 * w2
        Hp, Wp = (h2 + 6) // 7 * 7, (w2 + 6) // 7 * 7
        nW = (Hp // 7) * (Wp // 7)
        x32 = ws.get("x32", (max(n_loc, 1) * HW, E), _F, device=dev)
        c16 = ws.get("c16", (max(n_loc, 1) * HW, E), _H, device=dev)
        if n_loc:
            ops.head_fuse(proj, sizes, n_loc, E, 0, P["shift"], half32=x32, half16=c16)
        # ---- reference frames: K/V of every block, packed for the exchange (cffm_transformer.py:780-805, :495-518).  The local
        # frames are frame-major, so the reference frames of one temporal role are contiguous: one pooling launch per role and
        # ONE K/V projection per block over the rank's packed [role 0 maps | role 1 maps | role 2 maps] token buffer.
        role_off, flat_tok = plan.role_offsets(nW)
        send = ws.get("send", (depth, flat_tok, 2 * E), _H, device=dev, zero=True)
        # ---- the exchange, block by block, on a side stream beside the target frames' path: the K/V of block i
        # are gathered as soon as they exist, the CFM launch of block i waits for ITS gather only -- the reference-frame work
        # and the all-gather of block i+1 run under block i's attention and FFN.  The CFM kernel reads the gathered buffer IN
        # PLACE through a per-clip slot table.
        if plan.G > 1:
            gathered = ws.get("gathered", (depth, plan.G, flat_tok, 2 * E), _H, device=dev)
        else:
            gathered = send.view(depth, 1, flat_tok, 2 * E)
        ready = []
        with ops.fork("gather"):
            if n_ref:
                xn = ws.get("xn_ref", (n_ref * HW, E), _H, device=dev)
                pooled = ws.get("pooled_refs", (flat_tok, E), _H, device=dev, zero=True)     # rows of unused slots stay zero
            for i, b in enumerate(P["blocks"]):
                if n_ref:
                    ops.cffa_norm_frames(x32[:n_ref * HW], b["n1g"], b["n1b"], b["n1eps"], xn, None, n_ref, n_ref, h2, w2, Hp, Wp, E)
                    first = 0
                    for k, per in enumerate(ROLE_TOKENS):
                        n_k = plan.role_count[self.rank][k]
                        if n_k:
                            ops.cffa_pool_level(xn[first * HW:(first + n_k) * HW], n_k, k + 1, h2, w2, E, b["pool_w"], b["pool_b"],
                                                pooled[role_off[k]:role_off[k] + n_k * per * nW])
                        first += n_k
                    ops.gemm(pooled, b["qkv_w"][E:], bias=b["qkv_b"][E:], out16=send[i])
                if plan.G > 1:
                    all_gather_slots(send[i], self.group, out=gathered[i].view(plan.G * flat_tok, 2 * E))
                ev = torch.cuda.Event()
                ev.record()                                      # on the side stream: block i's reference K/V are in place
                ready.append(ev)
        if not n_t:
            ops.join("gather")
            return torch.empty(0, H, W, dtype=torch.int64, device=dev)
        # ---- target frames: CFFM blocks (cffm_transformer.py:709-832) with the gathered reference K/V
        xt = x32[n_ref * HW:(n_ref + n_t) * HW]
        ct16 = c16[n_ref * HW:(n_ref + n_t) * HW]
        xn_t = ws.get("xn_t", (n_t * HW, E), _H, device=dev)
        xt_pad = ws.get("xt_pad", (ops.apron_rows(n_t, h2, w2), E), _H, device=dev)
        pooled_t = ws.get("pooled_t", (n_t * nW, E), _H, device=dev)
        kv_t = ws.get("kv_t", (n_t * nW, 2 * E), _H, device=dev)
        qkv_t = ws.get("qkv_t", (ops.apron_rows(n_t, h2, w2), 3 * E), _H, device=dev)
        ao = ws.get("ao", (n_t * HW, E), _H, device=dev)
        xn2 = ws.get("xn2", (n_t * HW, E), _H, device=dev)
        hid = ws.get("hid", (n_t * HW, 4 * E), _H, device=dev)
        xt16 = ws.get("xt16", (n_t * HW, E), _H, device=dev)
        slots = self._slot_table(dev)
        for i, b in enumerate(P["blocks"]):
            ops.cffa_norm_frames(xt, b["n1g"], b["n1b"], b["n1eps"], xn_t, xt_pad, n_t, 0, h2, w2, Hp, Wp, E)
            ops.cffa_pool_level(xn_t, n_t, 0, h2, w2, E, b["pool_w"], b["pool_b"], pooled_t)
            ops.gemm(xt_pad, b["qkv_w"], bias=b["qkv_b"], out16=qkv_t)
            ops.gemm(pooled_t, b["qkv_w"][E:], bias=b["qkv_b"][E:], out16=kv_t)
            torch.cuda.current_stream().wait_event(ready[i])
            ops.cfm_attention_slots(qkv_t, kv_t, gathered[i], role_off, plan.role_slots, slots, b["bias"], ao, n_t, h2, w2, E, HEADS,
                                    (E // HEADS) ** -0.5)
            ops.gemm(ao, b["proj_w"], bias=b["proj_b"], residual=xt, out32=xt)
            ops.layernorm(xt, b["n2g"], b["n2b"], b["n2eps"], out16=xn2)
            ops.gemm(xn2, b["f1w"], bias=b["f1b"], out16=hid, act=ops.ACT_GELU)
            ops.gemm(hid, b["f2w"], bias=b["f2b"], residual=xt, out32=xt, out16=xt16 if i == depth - 1 else None)
        ops.join("gather")
        # ---- classifier + fused x2 / x4 resize + argmax (cffm_head.py:145-149, encoder_decoder.py:373-377,564)
        lg = ws.get("lg", (n_t * HW, P["ncp"]), _F, device=dev)
        ops.gemm(ct16, P["pred2_w"][:, :E], bias=P["pred2_b"], out32=lg)
        ops.gemm(xt16, P["pred2_w"][:, E:], residual=lg, out32=lg)
        labels = torch.empty(n_t, H, W, dtype=torch.int64, device=dev)
        if ops.upsample2_argmax_supported(h, w, H, W):
            ops.upsample2_argmax(lg, model.num_classes, labels, n_t, h2, w2, h, w, H, W)
        else:
            logits = torch.empty(n_t, model.num_classes, h, w, dtype=_F, device=dev)
            ops.resize_nhwc_to_nchw(lg, model.num_classes, logits, n_t, h2, w2, h, w)
            ops.resize_argmax(logits, labels, n_t, model.num_classes, h, w, H, W)
        return labels


class GraphedFrameShard:
    """One frame-sharded pass (kernels + the NCCL all-gather) captured in a CUDA graph and replayed with a single launch per
    rank.  NCCL collectives are capturable once the communicator exists, hence the eager warm-up runs first; every rank must
    capture and replay in lock step."""

    def __init__(self, runner, frames, warmup=2):
        self.runner = runner
        self.frames = frames                                     # static input buffer: copy new frames into it, then replay()
        dev = frames.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                runner.run(self.frames)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _abi.n_launches
        with torch.cuda.graph(self.graph):
            self.labels = runner.run(self.frames)
        self.kernels_per_replay = _abi.n_launches - n0
        for ws in [runner.ws] + runner.model.workspaces():       # the graph replays on these addresses: never free them
            ws.pin()

    def close(self):
        """Release the captured graph (it holds the communicator's resources: destroy_process_group() hangs while a graph
        with a captured collective is alive)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def replay(self):
        self.graph.replay()
        _abi.n_launches += self.kernels_per_replay
        return self.labels
