"""Bring-up tool: SM-clock timeline of the CFM kernel's pipeline events (cffm_cfm_attention_prof), B=2, 60x60."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vss_cffm_b200 import _abi, ops
from vss_cffm_b200 import cffm_tables as tb

torch.set_grad_enabled(False)
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
B, H, W, E, nW = 2, 60, 60, 256, 81
lay = ops.cfm_layout()
qkv, kvp = rn(ops.apron_rows(B, H, W), 3 * E).half(), rn(B * 15 * nW, 2 * E).half()
bias = tb.assemble_bias_tc(rn(169, 8) * 0.1, rn(1, 8, 49, 132) * 0.1, rn(8, 121) * 0.1, [rn(8, 169) * 0.1, rn(8, 121) * 0.1, rn(8, 81) * 0.1], 32 ** -0.5, lay)
out = torch.empty(B * H * W, E, device="cuda", dtype=torch.half)
lib = _abi.load()
fn = lib.cffm_cfm_attention_prof
vp, i32 = ctypes.c_void_p, ctypes.c_int
fn.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, ctypes.c_float, vp]
fn.restype = i32
grid = 148
prof = torch.zeros(grid, 8, 32, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.fill_(1)
    prof.zero_()
    st = fn(qkv.data_ptr(), kvp.data_ptr(), bias.data_ptr(), out.data_ptr(), prof.data_ptr(), B, H, W, E, 8, 32 ** -0.5,
            torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.cffm_last_error()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush.fill_(1)
ev0.record()
fn(qkv.data_ptr(), kvp.data_ptr(), bias.data_ptr(), out.data_ptr(), prof.data_ptr(), B, H, W, E, 8, 32 ** -0.5, torch.cuda.current_stream().cuda_stream)
ev1.record()
torch.cuda.synchronize()
print(f"last launch: {ev0.elapsed_time(ev1) * 1e3:.1f} us (CUDA events)")
p = prof.cpu()
names = ["prodK_issue", "prodQ_done", "prodV_issue", "mma_kfull", "mma_qfull", "mma_sempty", "mma_p0full", "mma_vfull", "sm_start",
         "sm_sfull", "sm_pass1", "sm_pass2", "sm_ofull", "mma_item_done", "p1c0", "p1c1", "p1c2", "ep_ofull", "ep_ld", "ep_stg", "p2c1_ld", "p2c1_math",
         "p2c1_pempty", "p2c1_st", "p2c1_arrive", "p2c2_ld"]
g0, g1 = p[:, 7, 26], p[:, 7, 27]
cyc = (p[:, 7, 31] - p[:, 7, 28]).double()
ns = (g1 - g0).double()
print(f"kernel span by globaltimer: {(g1.max() - g0.min()).item() / 1e3:.2f} us; per-CTA ns {ns.min().item():.0f}..{ns.max().item():.0f}; "
      f"per-CTA cycles {cyc.min().item():.0f}..{cyc.max().item():.0f}; implied SM clock {1e3 * (cyc / ns).median().item():.0f} MHz")
for cta in (0, 77):
    t0 = p[cta, :7][p[cta, :7] > 0].min().item()
    print(f"CTA {cta}: entry={p[cta,7,28].item()-t0} prologue_done={p[cta,7,29].item()-t0} after_pdl={p[cta,7,30].item()-t0} exit={p[cta,7,31].item()-t0}")
    print(f"--- CTA {cta} (cycles since first event)")
    for it in range(1, 3):
        row = p[cta, it]
        if row.max() == 0:
            break
        print(f" item {it}: " + "  ".join(f"{n}={row[i].item() - t0}" for i, n in enumerate(names) if row[i] > 0))
