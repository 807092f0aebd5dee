"""torchrun --nproc-per-node G tools/check_frame_shard.py [clips]: frame-sharded inference (one NCCL all-gather of the
reference-frame K/V) must reproduce the single-GPU labels bit for bit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import vss_cffm_b200 as V
from vss_cffm_b200 import parallel, synth

torch.set_grad_enabled(False)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
Bg = int(sys.argv[1]) if len(sys.argv) > 1 else 2 * world
H = W = 480
assert V._abi.load().cffm_current_device() == int(os.environ["LOCAL_RANK"])
m = V.build_segmentor(V.model_cfg("b1"))
synth.fill_module(m, 21)
m = m.cuda().eval()
gen = lambda b, t: synth.synth_array((3, H, W), 7000 + 16 * b + t)
plan = parallel.FrameShardPlan(Bg, 4, world)
runner = parallel.FrameShardedRunner(m, plan, rank)
frames = torch.stack([gen(b, t) for b, t in runner.local_frames()]).cuda()
got = runner.run(frames)
torch.cuda.synchronize()
ok = True
for j, clip in enumerate(plan.targets[rank]):
    imgs = [gen(clip, t).unsqueeze(0).cuda() for t in range(4)]
    ref = m.predict_labels(imgs, synth.img_metas(1, H, W))[0]
    same = torch.equal(ref, got[j])
    print(f"rank {rank}: clip {clip}: frame-sharded == single-GPU: {same} (agreement {(ref == got[j]).float().mean().item():.6f})", flush=True)
    ok &= same
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("FRAME_SHARD_CHECK", "PASS" if flag.item() else "FAIL", f"world={world} clips={Bg} frames/rank={[len(f) for f in plan.frames]}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
