"""Workspace buffer lifetime: a captured CUDA graph bakes buffer addresses in, so a pinned workspace never frees one."""
import torch

from vss_cffm_b200.workspace import Workspace


def test_unpinned_workspace_replaces_on_shape_change():
    ws = Workspace()
    a = ws.get("x", (4, 8), torch.float32, device="cpu")
    assert ws.get("x", (4, 8), torch.float32, device="cpu") is a
    b = ws.get("x", (2, 8), torch.float32, device="cpu")
    assert b is not a and ws.nbytes() == 2 * 8 * 4             # the old buffer is gone


def test_pinned_workspace_never_frees_and_reuses_parked_buffers():
    ws = Workspace()
    a = ws.get("x", (4, 8), torch.float32, device="cpu", zero=True)
    ws.pin()                                                    # a graph was captured over `a`
    b = ws.get("x", (2, 8), torch.float32, device="cpu")       # eager call with another geometry
    assert b is not a and b.data_ptr() != a.data_ptr()
    assert ws.nbytes() == (4 * 8 + 2 * 8) * 4                   # `a` is parked, still alive
    a2 = ws.get("x", (4, 8), torch.float32, device="cpu")      # back to the captured geometry: the very same memory
    assert a2 is a
    assert ws.get("x", (2, 8), torch.float32, device="cpu") is b   # and no growth from flipping back and forth
    assert ws.nbytes() == (4 * 8 + 2 * 8) * 4
    ws.clear()
    assert ws.nbytes() == 0 and not ws.pinned
