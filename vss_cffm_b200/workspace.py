"""Named, shape-keyed device buffers.  PyTorch stays the allocator (the C ABI
never allocates); buffers persist across calls so that a forward pass has stable addresses and can
be captured in a CUDA graph."""
import torch


class Workspace:
    """``get(name, shape, dtype)`` returns the same tensor for the same request.  Until the workspace is pinned a
    request with a new shape REPLACES the buffer of that name (the old memory is freed).  A captured CUDA graph bakes
    buffer addresses in, so whoever captures a graph over this workspace calls ``pin()``: from then on a buffer is
    never freed -- a request with another shape gets its own tensor, the old one is parked and handed out again when
    its exact (shape, dtype, device) is asked for -- and an eager call with a different geometry can no longer
    invalidate the memory a graph replays on.  Memory then grows with the number of distinct geometries used."""

    def __init__(self):
        self._bufs = {}
        self._parked = {}
        self._pins = 0

    def get(self, name, shape, dtype, device=None, zero=False):
        """Buffer ``name`` of exactly ``shape``.  ``zero=True`` zero-fills on allocation only."""
        shape = tuple(int(s) for s in shape)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype or t.device != device:
            if t is not None and self._pins:
                self._parked[(name, tuple(t.shape), t.dtype, t.device)] = t
            t = self._parked.pop((name, shape, dtype, device), None)
            if t is None:
                t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            self._bufs[name] = t
        return t

    def pin(self):
        """Called by the owner of a CUDA graph captured over these buffers."""
        self._pins += 1

    @property
    def pinned(self):
        return self._pins > 0

    def clear(self):
        """Drops every buffer.  Graphs captured over them must be discarded first."""
        self._bufs.clear()
        self._parked.clear()
        self._pins = 0

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in list(self._bufs.values()) + list(self._parked.values()))
