"""Bring-up tool: SM-clock timeline of the MHA kernel's pipeline events (cffm_mha_f16_prof), MiT stage-1 shape."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vss_cffm_b200 import _abi

torch.set_grad_enabled(False)
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
Nf, Nq, Nkv, heads, d = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (8, 14400, 225, 1, 64)
C = heads * d
q, kv, out = rn(Nf * Nq, C).half(), rn(Nf * Nkv, 2 * C).half(), torch.empty(Nf * Nq, C, device="cuda", dtype=torch.half)
lib = _abi.load()
fn = lib.cffm_mha_f16_prof
vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
fn.argtypes = [vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, ctypes.c_float, vp, vp]
fn.restype = i32
grid = 148
prof = torch.zeros(grid, 8, 32, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
k, v = kv[:, :C], kv[:, C:]
call = lambda: fn(q.data_ptr(), C, k.data_ptr(), v.data_ptr(), 2 * C, out.data_ptr(), C, Nf, Nq, Nkv, heads, d, d ** -0.5,
                  prof.data_ptr(), torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    flush.fill_(1)
    prof.zero_()
    st = call()
    assert st == 0, lib.cffm_last_error()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush.fill_(1)
ev0.record()
call()
ev1.record()
torch.cuda.synchronize()
print(f"last launch: {ev0.elapsed_time(ev1) * 1e3:.1f} us (CUDA events)")
p = prof.cpu()
names = ["prod_kvempty", "prod_qempty", "-", "qk_qfull", "qk_sempty0", "qk_sempty1", "pv_p0full", "pv_done", "sm_start", "sm_sfull",
         "sm_pass1", "sm_pass2", "sm_sum", "sm_ofull", "sm_done"]
g0, g1 = p[:, 7, 26], p[:, 7, 27]
cyc = (p[:, 7, 31] - p[:, 7, 28]).double()
ns = (g1 - g0).double()
print(f"kernel span by globaltimer: {(g1.max() - g0.min()).item() / 1e3:.2f} us; per-CTA ns {ns.min().item():.0f}..{ns.max().item():.0f}; "
      f"per-CTA cycles {cyc.min().item():.0f}..{cyc.max().item():.0f}; implied SM clock {1e3 * (cyc / ns).median().item():.0f} MHz")
for cta in (0, 77):
    t0 = p[cta, 7, 28].item()
    print(f"CTA {cta}: prologue_done={p[cta,7,29].item()-t0} after_pdl={p[cta,7,30].item()-t0} exit={p[cta,7,31].item()-t0}")
    for it in range(0, 7):
        row = p[cta, it]
        if row[:15].max() == 0:
            break
        print(f" item {it}: " + "  ".join(f"{n}={row[i].item() - t0}" for i, n in enumerate(names) if row[i] > 0))
