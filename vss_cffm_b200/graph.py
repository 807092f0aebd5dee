"""CUDA-graph replay of one EncoderDecoder_clips inference pass (fixed shapes).

The forward is a fixed sequence of C-ABI kernel launches on stable Workspace addresses with no host
synchronisation, so it is captured once (torch.cuda.CUDAGraph only provides the capture stream and the
private memory pool for the output tensor) and then replayed with a single cudaGraphLaunch per step."""
import torch

from . import _abi


class GraphedClips:
    def __init__(self, model, B, T, H, W, img_meta=None, rescale=True, head_kw=None, warmup=3, private_input=False,
                 preprocessor=None, src_hw=None, private_workspace=False):
        """``preprocessor`` (a ``ClipPreprocessor``) + ``src_hw``: the captured pass starts from a static uint8 BGR HWC
        buffer ``self.frames_u8`` (T, B, h, w, 3) of decoded frames instead of the normalised fp32 frames.
        ``private_workspace``: the pass is captured over its OWN intermediate buffers (the model's workspaces are swapped
        for fresh ones during warm-up and capture), so two such graphs may replay CONCURRENTLY on different streams; the
        weights (plans) stay shared, they are read-only."""
        dev = model._device()
        if dev.type != "cuda":
            raise _abi.CffmError("CUDA graphs need the model on a CUDA device")
        self.model, self.B, self.T = model, B, T
        self.meta = img_meta or [dict(ori_shape=(H, W, 3), img_shape=(H, W, 3), pad_shape=(H, W, 3), flip=False,
                                      filename="data/video/origin/00000000.jpg") for _ in range(B)]
        head_kw = dict(head_kw or {})
        if private_input:                                        # pipeline slots: one static input buffer per graph
            self.frames = torch.empty(T, B, 3, H, W, dtype=torch.float32, device=dev)
        else:
            self.frames = model._ws.get("frames", (T, B, 3, H, W), torch.float32, device=dev)   # static input buffer
        self.frames_u8 = None
        if preprocessor is not None:
            h, w = src_hw
            assert preprocessor.output_size(h, w)[1] == (H, W), "preprocessor output size != model input size"
            self.frames_u8 = torch.zeros(T, B, h, w, 3, dtype=torch.uint8, device=dev)

        def run_pass():
            if preprocessor is not None:
                preprocessor.run(self.frames_u8.view(T * B, *self.frames_u8.shape[2:]), T, B, out=self.frames)
            return model.labels_from_frames(self.frames, self.meta, rescale, **head_kw)

        saved = []
        if private_workspace:
            from .workspace import Workspace
            for owner in (model, model.backbone, model.decode_head):
                if hasattr(owner, "_ws"):
                    saved.append((owner, owner._ws))
                    owner._ws = Workspace()
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                        # plans, workspaces, func attributes: all set here
                for _ in range(warmup):
                    run_pass()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            n0 = _abi.n_launches
            with torch.cuda.graph(self.graph):
                self.labels = run_pass()
            self.kernels_per_replay = _abi.n_launches - n0
            self._workspaces = list(model.workspaces())
            for ws in self._workspaces:                          # the graph replays on these addresses: never free them
                ws.pin()
        finally:
            for owner, ws in saved:
                owner._ws = ws
        # ... and on the folded weights of the plans it was captured with: keep them alive even if a later
        # load_state_dict() rebuilds the plans (the owner is expected to drop this graph then: invalidate_graphs())
        self._plans = (getattr(model.backbone, "_plan", None), getattr(model.decode_head, "_plan", None))

    def load(self, imgs):
        """Copy a list of T (B,3,H,W) tensors (host, ideally pinned, or device) into the static input buffer."""
        assert len(imgs) == self.T
        for t, f in enumerate(imgs):
            self.frames[t].copy_(f, non_blocking=True)

    def replay(self):
        """Run the captured pass on whatever is in ``self.frames``; returns the static labels tensor (B,H,W)."""
        self.graph.replay()
        _abi.n_launches += self.kernels_per_replay
        return self.labels

    def __call__(self, imgs):
        self.load(imgs)
        return self.replay()


class ClipPipeline:
    """Streaming inference from HOST buffers: the H2D copy of clip batch i+1 and the D2H read of the labels of
    batch i-1 overlap the kernels of batch i.  ``depth`` graph instances (one static input buffer and one label
    buffer each, OWN workspace) rotate; the streams (copy-in, one compute stream PER SLOT, copy-out) are ordered with events
    only, so ``submit`` never blocks the host unless all slots are in flight.  Because every slot owns its intermediate
    buffers, the passes of consecutive batches also overlap ON the GPU: a pass is a dependent chain of ~110 kernels, most of
    them small, and the bubbles of one chain (launch-to-launch latency, tiles that fill a fraction of the SMs) are filled by
    the other (+19 % throughput with two passes in flight, tools/overlap_probe.py).

        pipe = ClipPipeline(model, B, T, H, W)
        tickets = [pipe.submit(frames_host[i], labels_host[i]) for i in ...]   # pinned host tensors
        pipe.wait(tickets[-1])                                                   # or pipe.drain()
    """

    def __init__(self, model, B, T, H, W, img_meta=None, rescale=True, head_kw=None, depth=2, preprocessor=None, src_hw=None):
        """With ``preprocessor`` / ``src_hw`` the pipeline is fed decoded uint8 BGR HWC frames (T, B, h, w, 3): a quarter of
        the H2D bytes, and the resize / normalise / layout change run on the GPU inside the captured pass."""
        dev = model._device()
        self.dev, self.depth, self.T = dev, depth, T
        self.slots = []
        for s in range(depth):
            g = GraphedClips(model, B, T, H, W, img_meta, rescale, head_kw, warmup=3 if s == 0 else 1, private_input=True,
                             preprocessor=preprocessor, src_hw=src_hw, private_workspace=True)
            self.slots.append(dict(g=g, h2d=torch.cuda.Event(), done=torch.cuda.Event(), d2h=torch.cuda.Event(), used=False,
                                   run=torch.cuda.Stream(device=dev)))
        self.s_in, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(2))
        self.n = 0

    @property
    def compute_stream(self):
        """Compute stream of the slot the NEXT submit will use."""
        return self.slots[self.n % self.depth]["run"]

    def submit(self, imgs, labels_host):
        """imgs: list of T (B,3,H,W) pinned host tensors (or one (T,B,3,H,W)); labels_host: pinned int64 (B,H,W)."""
        sl = self.slots[self.n % self.depth]
        s_run = sl["run"]
        g = sl["g"]
        with torch.cuda.stream(self.s_in):
            if sl["used"]:
                self.s_in.wait_event(sl["done"])                 # the previous batch in this slot has been consumed
            dst = g.frames if g.frames_u8 is None else g.frames_u8
            if isinstance(imgs, (list, tuple)):
                for t, f in enumerate(imgs):
                    dst[t].copy_(f, non_blocking=True)
            else:
                dst.copy_(imgs, non_blocking=True)
            sl["h2d"].record(self.s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(sl["h2d"])
            if sl["used"]:
                s_run.wait_event(sl["d2h"])                      # its label buffer has been read back
            g.replay()
            sl["done"].record(s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(sl["done"])
            labels_host.copy_(g.labels, non_blocking=True)
            sl["d2h"].record(self.s_out)
        sl["used"] = True
        self.n += 1
        return sl["d2h"]

    @staticmethod
    def wait(ticket):
        ticket.synchronize()

    def drain(self):
        for sl in self.slots:
            if sl["used"]:
                sl["d2h"].synchronize()

    @property
    def kernels_per_step(self):
        return self.slots[0]["g"].kernels_per_replay

