"""Host logic of the multi-GPU path on CPU: ownership plan, slot packing and the ONE all-gather, exercised with
world_size 2 over gloo (127.0.0.1).  The payload carries (clip, role, block) codes instead of real K/V so that every
rank can check it assembled exactly the reference frames of its own target clips."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vss_cffm_b200.parallel import ROLE_TOKENS, FrameShardPlan, all_gather_slots, assemble_kv, shard_clips


def test_shard_clips_partitions_without_overlap():
    for n, g in ((16, 8), (7, 3), (2, 4), (5, 1)):
        seen = []
        for r in range(g):
            seen += list(shard_clips(n, g, r))
        assert seen == list(range(n))


@pytest.mark.parametrize("Bg,G", [(16, 8), (2, 2), (4, 2), (2, 8), (8, 4), (3, 2)])
def test_plan_covers_every_frame_once(Bg, G):
    plan = FrameShardPlan(Bg, 4, G)
    flat = [f for fr in plan.frames for f in fr]
    assert sorted(flat) == sorted((b, t) for b in range(Bg) for t in range(4))
    for r in range(G):
        assert plan.frames[r] == sorted(plan.frames[r], key=lambda f: (f[1], f[0]))          # frame-major
        assert plan.frames[r][:len(plan.refs[r])] == plan.refs[r]                             # refs first, targets last
        for i, f in enumerate(plan.refs[r]):
            assert plan.slot_of(*f) == (r, i) and plan.gathered_index(*f) == r * plan.max_slots + i
    assert sorted(b for t in plan.targets for b in t) == list(range(Bg))
    if Bg % G == 0:                                                                           # balanced: Bg/G per role
        assert all(len(fr) == 4 * Bg // G and len(tg) == Bg // G for fr, tg in zip(plan.frames, plan.targets))
    with pytest.raises(Exception):
        FrameShardPlan(2, 3, 2)


def _code(b, t, block, tok):
    return float(((b * 4 + t) * 8 + block) * 4096 + tok)


def _worker(rank, world, port, Bg, depth, nW, errs):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        plan = FrameShardPlan(Bg, 4, world)
        send = torch.zeros(plan.max_slots, depth, 9 * nW, 4)
        for slot, (b, t) in enumerate(plan.refs[rank]):
            for blk in range(depth):
                n = ROLE_TOKENS[t] * nW
                send[slot, blk, :n, 0] = torch.tensor([_code(b, t, blk, k) for k in range(n)])
        gathered = all_gather_slots(send)
        assert gathered.shape[0] == world * plan.max_slots
        for clip in plan.targets[rank]:
            for blk in range(depth):
                out = torch.full((15 * nW, 4), -1.0)
                assemble_kv(plan, gathered, clip, blk, nW, out)
                assert (out[:nW] == -1).all()                               # pooled-target rows are not touched
                off = nW
                for t, per in enumerate(ROLE_TOKENS):
                    n = per * nW
                    want = torch.tensor([_code(clip, t, blk, k) for k in range(n)])
                    assert torch.equal(out[off:off + n, 0], want), (rank, clip, blk, t)
                    off += n
                assert off == 15 * nW
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                                   # pragma: no cover
        errs.put(f"rank {rank}: {type(e).__name__}: {e}")
        raise


@pytest.mark.parametrize("Bg", [2, 4])
def test_kv_all_gather_world2_gloo(Bg):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    errs = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, Bg, 2, 6, errs)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    msgs = []
    while not errs.empty():
        msgs.append(errs.get())
    assert not msgs, msgs
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


@pytest.mark.parametrize("Bg,G", [(16, 8), (4, 2), (2, 2), (3, 2), (2, 8)])
def test_role_packing_indexes_every_reference_frame(Bg, G):
    """The packed per-role layout of the exchange: (index, rank) of every reference frame is unique inside its role and below
    the role's slot count; offsets tile the buffer without overlap."""
    plan = FrameShardPlan(Bg, 4, G)
    nW = 81
    off, total = plan.role_offsets(nW)
    assert off[0] == 0 and total == sum(plan.role_slots[k] * ROLE_TOKENS[k] * nW for k in range(3))
    for k in range(3):
        seen = set()
        for b in range(Bg):
            idx, rank = plan.role_index(b, k)
            assert rank == plan.owner(b, k) and 0 <= idx < plan.role_slots[k] and (idx, rank) not in seen
            seen.add((idx, rank))
        assert plan.role_slots[k] == max(1, max(c[k] for c in plan.role_count))
    # the frames of a rank are frame-major, so its role-k frames are contiguous and in role_index order
    for r in range(G):
        pos = {0: 0, 1: 0, 2: 0}
        for (b, t) in plan.refs[r]:
            assert plan.role_index(b, t) == (pos[t], r)
            pos[t] += 1
