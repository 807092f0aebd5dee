"""Named, shape-keyed device buffers.  PyTorch stays the allocator (the C ABI
never allocates); buffers persist across calls so that a forward pass has stable addresses and can
be captured in a CUDA graph."""
import torch


class Workspace:
    def __init__(self):
        self._bufs = {}

    def get(self, name, shape, dtype, device=None, zero=False):
        """Buffer ``name`` of exactly ``shape``; reallocated (never resized in place) on a shape change.
        ``zero=True`` zero-fills on (re)allocation only."""
        shape = tuple(int(s) for s in shape)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype or t.device != device:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            self._bufs[name] = t
        return t

    def clear(self):
        self._bufs.clear()

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self._bufs.values())
