#!/usr/bin/env python
"""bench.py -- clip-frames/s of the CFFM hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): MiT-B1 + CFFM head (depth 2), 480x480, T=4, batch 2 clips per GPU,
synthetic N(0,1) frames, synthetic weights (vss_cffm_b200.synth), 124 classes.  One "step" = one
EncoderDecoder_clips inference pass over one batch: frames -> int64 label maps.

  value : clip-frames/s (B*T*N / max-over-ranks device time), inputs resident in HBM, labels left in HBM
  e2e   : same metric through the public API with HOST (pinned) DECODED frames (uint8): H2D copy, preprocessing,
          forward, D2H of the labels inside the timed region (e2e_from_fp32_tensors: fed normalised fp32 tensors;
          e2e_mmseg_call: the reference-facing model(img=..., return_loss=False) call itself)
  roofline     : the dominant kernel family (tcgen05 GEMM) and, separately, the CFM attention kernel the metric
                 names, both timed live with CUDA events around each launch
  cpu_baseline : the CPU oracle (a port of the reference's PyTorch path) on this box's host cores, bounded sample

N>1: clips shard across ranks (the reference's own data-parallel strategy): weak
scaling, no data-path collective; `--shard frames` exercises the frame-sharded path with one NCCL
all-gather of the reference-frame features.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clip-frames/sec (480x480, T=4, MiT-B1+CFFM)"
UNIT = "clip-frames/s"
H = W = 480
T = 4
CLIPS_PER_GPU = 2
VARIANT = "b1"
HEAD_DEPTH = {"b0": 1, "b1": 2, "b2": 2, "b5": 4}                   # local_configs/cffm/B*/: decoder_params.depths
KIND = "cffm"
PROTOS = 0
CFM_FLOPS_PER_CLIP_BLOCK = 2 * 2 * 81 * 8 * 49 * 289 * 32            # QK^T + PV, SURVEY.md 8(d): 1.1746 GFLOP
# fp16 operands each once, per clip per block: Q + target K,V + pooled K,V + O + bias tables
CFM_BYTES_PER_CLIP_BLOCK = (3969 * 256 * 2) + (3969 * 512 * 2) + (1215 * 512 * 2) + (3600 * 256 * 2) + (8 * 49 * 289 * 4)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shard", default="clips", choices=["clips", "frames"])
    ap.add_argument("--variant", default="b1", choices=["b0", "b1", "b2", "b5"],
                    help="MiT backbone; b1 is the headline workload (BASELINE configs[1]), b2 = configs[3], b0 with --clips 1 --T 2 = configs[0]")
    ap.add_argument("--clips", type=int, default=2, help="clips per GPU and step (headline: 2)")
    ap.add_argument("--T", type=int, default=4, help="frames per clip; T != 4 takes the head's early-return path (configs[0]: T = 2)")
    ap.add_argument("--kind", default="cffm", choices=["cffm", "cffmpp"], help="cffmpp = CFFM++ head with k-means prototypes (configs[4])")
    ap.add_argument("--protos", type=int, default=64, help="prototypes per clip for --kind cffmpp (configs[4]: 64; README: 100)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def oracle_state(seed=21):
    """Weights for the CPU oracle: the same synthetic state dict the GPU model is filled with."""
    import vss_cffm_b200 as V
    from vss_cffm_b200 import synth
    m = V.build_segmentor(V.model_cfg(VARIANT, KIND))
    synth.fill_module(m, seed)
    return m, {k: v.clone() for k, v in m.state_dict().items()}


def cpu_reference_time(sd, steps, warmup, clips=1):
    """Times the CPU oracle (oracle/cffm_oracle.py: plain fp32 PyTorch restatement of the reference's
    EncoderDecoder_clips forward) on `clips` clips per step with every host thread torch will use."""
    import torch
    from oracle import cffm_oracle as O
    from vss_cffm_b200 import synth
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    imgs = synth.synth_clip(clips, T, H, W, seed=3)
    centers = synth.synth_array((clips, PROTOS, 256), 77) if KIND == "cffmpp" else None
    for _ in range(warmup):
        O.segmentor_simple_test(sd, imgs, "mit_" + VARIANT, HEAD_DEPTH[VARIANT], centers=centers)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.segmentor_simple_test(sd, imgs, "mit_" + VARIANT, HEAD_DEPTH[VARIANT], centers=centers).numpy()
        ts.append(time.perf_counter() - t0)
    mean = sum(ts) / len(ts)
    return clips * T / mean, mean * 1e3, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    if os.environ.get("OMP_NUM_THREADS") and not os.environ.get("CFFM_REF_CHILD"):
        # torchrun pins OMP_NUM_THREADS=1 before the interpreter starts: the OpenMP pool of THIS process stays at one thread
        # whatever torch.set_num_threads() says later.  The CPU arm must use every host core: re-run it in a clean child.
        import subprocess
        drop = ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK",
                "ROLE_RANK", "ROLE_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")
        env = {k: v for k, v in os.environ.items() if k not in drop and not k.startswith("TORCHELASTIC")}
        env["CFFM_REF_CHILD"] = "1"
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--variant", args.variant, "--clips", str(args.clips), "--T", str(args.T),
               "--kind", args.kind, "--protos", str(args.protos)]
        out = subprocess.run(cmd, env=env, capture_output=True, text=True)
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if out.returncode == 0 and lines:
            print(lines[-1])
            return
        print(f"reference child failed (rc {out.returncode}): {out.stderr[-400:]}; timing in-process", file=sys.stderr)
    _, sd = oracle_state()
    steps = max(1, min(args.steps, 8))                       # bounded: ~2 s per clip on 8 cores
    value, ms, cores = cpu_reference_time(sd, steps, max(1, min(args.warmup, 1)), clips=1)
    sample = f"1 clip (T={T}, {H}x{W}) per step, {steps} timed steps after 1 warm-up, fp32, {cores} torch threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MiT-{VARIANT.upper()} + CFFM{'++ (K=%d)' % PROTOS if KIND == 'cffmpp' else ''} (depth {HEAD_DEPTH[VARIANT]}), {H}x{W}, T={T}, reference's PyTorch path on host CPU cores",
                   "note": "the reference is pure Python/PyTorch+mmcv (mmcv absent offline); this is oracle/cffm_oracle.py, the "
                           "CPU port pinned to goldens generated from the unmodified reference"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def measure_frame_shard(args, model, rank, world, B, flush, clip_ms):
    """Frame-sharded pass over B*world clips (vss_cffm_b200/parallel.py): bit-identity with the single-GPU path, throughput of
    the CUDA-graph replay (kernels + the all-gather), and the collective alone.  Returns the record (same on every rank)."""
    import torch
    import torch.distributed as dist
    from vss_cffm_b200 import parallel, synth
    plan = parallel.FrameShardPlan(B * world, T, world)
    runner = parallel.FrameShardedRunner(model, plan, rank)
    gen = lambda b, t: synth.synth_array((3, H, W), 7000 + 16 * b + t)
    fr_dev = torch.stack([gen(b, t) for b, t in runner.local_frames()]).cuda()
    got = runner.run(fr_dev).clone()
    torch.cuda.synchronize()
    same = True
    for j, clip in enumerate(plan.targets[rank]):                # the same clip through the single-GPU path of this rank
        ref = model.predict_labels([gen(clip, t).unsqueeze(0).cuda() for t in range(T)], synth.img_metas(1, H, W))[0]
        same &= bool(torch.equal(ref, got[j]))
    nodes, mode = None, "eager"
    step = lambda: runner.run(fr_dev)
    g = None
    try:
        g = parallel.GraphedFrameShard(runner, fr_dev)
        step, nodes, mode = g.replay, g.kernels_per_replay, "CUDA graph replay"
    except Exception as e:                                       # capture of the collective refused: stay eager, say so
        print(f"rank {rank}: frame-sharded pass not captured ({type(e).__name__}: {e}); launching eagerly", file=sys.stderr)
    for _ in range(3):
        step()
    dist.barrier(); torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    # the collective alone, on the payload of one step
    head = model.decode_head
    depth = len(head._plan["blocks"])
    nW = ((H // 8 + 6) // 7) * ((W // 8 + 6) // 7)
    send = torch.zeros(plan.role_offsets(nW)[1], 2 * head.embed_dim, dtype=torch.float16, device="cuda")   # one block's payload
    recv = torch.empty(world * send.shape[0], send.shape[1], dtype=torch.float16, device="cuda")
    for _ in range(3):
        parallel.all_gather_slots(send, out=recv)
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        parallel.all_gather_slots(send, out=recv)
    b.record()
    torch.cuda.synchronize()
    ag_us = a.elapsed_time(b) * 100.0
    if g is not None:
        g.close()
    t = torch.tensor([ms, ag_us, 0.0 if same else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ag_us, bad = t.tolist()
    frames = B * T * world * args.steps
    value = frames / (ms * 1e-3)
    clip_value = frames / (clip_ms * 1e-3)
    return {"value": round(value, 2), "unit": UNIT, "ms_per_step": round(ms / args.steps, 4),
            "workload": f"{B * world} clips (T={T}, {H}x{W}) with their {B * world * T} FRAMES spread over {world} GPUs "
                        f"({B * T} frames per GPU)" + (" (BASELINE configs[2])" if B * world == 16 and world == 8 else ""),
            "bit_identical_to_single_gpu": bad == 0.0, "vs_clip_sharded": round(value / clip_value, 4),
            "collective": f"{depth} ncclAllGather per step (one per CFFM block: the reference-frame K/V of block i) over NVLink / "
                          "NVSwitch, on a side stream together with the reference frames' norm1 / pooling / K,V projection; the CFM "
                          "launch of block i waits for ITS gather only, so block i+1's exchange runs under block i's attention and "
                          "FFN; the CFM kernel reads the gathered buffer in place",
            "all_gather_us": round(ag_us, 1), "all_gathers_per_step": depth, "all_gather_bytes_per_rank": int(send.numel() * 2),
            "all_gather_bytes_total": int(send.numel() * 2 * world),
            "launch_mode": mode + (f" ({nodes} kernel nodes + {depth} all-gathers per step)" if nodes else ""),
            "limiter": "the payload is ~1.2 MB per clip and block: latency-bound.  What frame sharding still costs against whole "
                       "clips per rank is the reference-frame chain of block 0 in front of the first exchange, and an all-gather that "
                       "delivers every rank's K/V to every rank (a rank needs 6 of the 8 N frames it receives)"}


def main():
    args = parse()
    global VARIANT, METRIC, T, CLIPS_PER_GPU, KIND, PROTOS
    VARIANT, T, CLIPS_PER_GPU, KIND = args.variant, args.T, args.clips, args.kind
    PROTOS = args.protos if KIND == "cffmpp" else 0
    METRIC = f"clip-frames/sec (480x480, T={T}, MiT-{VARIANT.upper()}+CFFM{'++' if KIND == 'cffmpp' else ''})"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    import vss_cffm_b200 as V
    from vss_cffm_b200 import _abi, ops, synth
    torch.set_grad_enabled(False)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (sm_100) device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    _abi.require_device()
    assert _abi.load().cffm_current_device() == local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    n_gpus = world

    model, sd = oracle_state()
    model = model.cuda().eval()
    B = CLIPS_PER_GPU
    imgs_host = [t.pin_memory() for t in synth.synth_clip(B, T, H, W, seed=100 + rank)]
    imgs_dev = [t.cuda() for t in imgs_host]
    metas = synth.img_metas(B, H, W)
    labels_host = torch.empty(B, H, W, dtype=torch.int64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    # CFFM++ (configs[4]): the prototypes of each clip's video are an INPUT of the head (k-means centres loaded from
    # <save_path>/<video>/centers.pt in the reference, cffm_head.py:429-455); the bench feeds synthetic ones
    head_kw = dict(centers=synth.synth_array((B, PROTOS, 256), 77).cuda()) if KIND == "cffmpp" else {}
    graphed = None
    gfs = None
    frames_graph_nodes = None
    frames_mode = args.shard == "frames" and world > 1 and T == 4 and KIND == "cffm"
    if frames_mode:
        # the global batch (B clips per GPU x world) with its FRAMES spread over the ranks; one NCCL all-gather of the
        # reference-frame K/V per step (vss_cffm_b200/parallel.py).  Launched eagerly (the collective is not captured).
        from vss_cffm_b200 import parallel
        plan = parallel.FrameShardPlan(B * world, T, world)
        runner = parallel.FrameShardedRunner(model, plan, rank)
        gen = lambda b, t: synth.synth_array((3, H, W), 7000 + 16 * b + t)
        fr_host = torch.stack([gen(b, t) for b, t in runner.local_frames()]).pin_memory()
        fr_dev = fr_host.cuda()
        labels_host = torch.empty(len(plan.targets[rank]), H, W, dtype=torch.int64).pin_memory()
        step_eager = lambda: runner.run(fr_dev)
        step_dev = step_eager
        if not args.no_graph:
            try:                                                 # kernels + the all-gather in one CUDA graph per rank
                gfs = parallel.GraphedFrameShard(runner, fr_dev)
                step_dev = gfs.replay
                frames_graph_nodes = gfs.kernels_per_replay
            except Exception as e:                               # capture of the collective refused: stay eager, say so
                print(f"rank {rank}: frame-sharded pass not captured ({type(e).__name__}: {e}); launching eagerly", file=sys.stderr)
    else:
        step_eager = lambda: model.predict_labels(imgs_dev, metas, **head_kw)
        if args.no_graph:
            step_dev = step_eager
        else:
            graphed = model.make_graphed(B, T, H, W, metas, **head_kw)      # one cudaGraphLaunch per step
            graphed.load(imgs_dev)
            step_dev = graphed.replay

    def step_e2e():
        if frames_mode:
            fr_dev.copy_(fr_host, non_blocking=True)
            lab = step_dev()
            labels_host.copy_(lab, non_blocking=True)
            return lab
        if graphed is not None:
            lab = graphed(imgs_host)                              # H2D of the pinned frames + graph replay
        else:
            lab = model.predict_labels(imgs_host, metas, **head_kw)
        labels_host.copy_(lab, non_blocking=True)
        return lab

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    checks = {}

    def check(name, ok):
        """Bit-identity checks of the measured modes against the direct call: reported in the JSON line (a failed check must
        not cost the whole line, it is printed loudly instead)."""
        checks[name] = bool(ok)
        if not ok:
            print(f"rank {rank}: CHECK FAILED: {name}", file=sys.stderr)

    def timed(fn, steps):
        """Sum of per-step CUDA-event times; L2 is flushed between steps outside the events."""
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n0 = _abi.n_launches
    serial_ms = timed(step_dev, args.steps)                      # one pass at a time: per-step events, flush outside them
    launches = _abi.n_launches - n0
    barrier()
    total_ms, passes_in_flight = serial_ms, 1
    if graphed is not None and not frames_mode:
        # Headline: K steps as TWO passes in flight.  Each pass is one CUDA-graph replay over its OWN intermediate buffers on its
        # own stream (the weights are shared, read-only), so the kernels of consecutive batches interleave on the GPU: a pass is a
        # dependent chain of ~110 kernels, most of them small, and one chain's bubbles are filled by the other.  The L2 flush of
        # every step runs INSIDE the timed region, on the step's stream.
        try:
            from vss_cffm_b200.graph import GraphedClips
            slots = [GraphedClips(model, B, T, H, W, metas, True, head_kw, warmup=2, private_input=True, private_workspace=True)
                     for _ in range(2)]
            slots[0].load(imgs_dev)
            slots[1].load([t.cuda() for t in synth.synth_clip(B, T, H, W, seed=300 + rank)])
            lanes = [torch.cuda.Stream() for _ in slots]
            chk0 = slots[0].replay().clone()
            torch.cuda.synchronize()
            check("two_passes_in_flight_labels_equal_direct_call", torch.equal(chk0, model.predict_labels(imgs_dev, metas, **head_kw)))

            def pipelined_dev(steps):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for st in lanes:
                    st.wait_event(e0)
                for i in range(steps):
                    with torch.cuda.stream(lanes[i & 1]):
                        flush.fill_(1)
                        slots[i & 1].replay()
                for st in lanes:
                    torch.cuda.current_stream().wait_stream(st)
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1)

            pipelined_dev(4)
            barrier()
            n0 = _abi.n_launches
            total_ms, passes_in_flight = pipelined_dev(args.steps), 2
            launches = _abi.n_launches - n0
            barrier()
        except Exception as e:                                   # keep the single-pass number rather than lose the line
            if world > 1:                                        # ... but never leave the other ranks waiting at a barrier
                raise
            print(f"rank {rank}: two passes in flight not measured ({type(e).__name__}: {e}); reporting one pass at a time", file=sys.stderr)
            total_ms, passes_in_flight = serial_ms, 1
    clocks = sampler.stop() if sampler else None

    for _ in range(3):
        step_e2e()
    barrier()
    e2e_serial_ms = timed(step_e2e, args.steps)
    barrier()
    e2e_ms, e2e_api = e2e_serial_ms, "EncoderDecoder_clips.predict_labels(pinned host frames) + D2H of the int64 label maps"
    e2e_u8, e2e_u8_ms = None, 0.0
    if graphed is not None:
        # streaming API: the H2D of batch i+1 and the D2H of batch i-1 overlap the kernels of batch i.  Every step still
        # copies its own frames from pinned host memory and reads its own labels back; the L2 flush runs INSIDE the
        # timed region (on the compute stream), so the number is a lower bound of the throughput.
        from vss_cffm_b200.graph import ClipPipeline
        pipe = ClipPipeline(model, B, T, H, W, metas, head_kw=head_kw)
        hosts = [imgs_host] + [[t.pin_memory() for t in synth.synth_clip(B, T, H, W, seed=200 + 7 * i + rank)] for i in (1, 2)]
        labs = [torch.empty(B, H, W, dtype=torch.int64).pin_memory() for _ in range(3)]

        def pipelined(steps):
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(pipe.s_in)
            for i in range(steps):
                with torch.cuda.stream(pipe.compute_stream):
                    flush.fill_(1)
                pipe.submit(hosts[i % 3], labs[i % 3])
            ev1.record(pipe.s_out)
            pipe.drain()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1)

        pipelined(3)
        # the pipelined labels must be the labels of the plain call on the same frames
        chk = model.predict_labels(hosts[2], metas, **head_kw)
        torch.cuda.synchronize()
        check("pipeline_labels_equal_direct_call", torch.equal(chk.cpu(), labs[2]))
        barrier()
        e2e_ms = pipelined(args.steps)
        e2e_api = ("ClipPipeline.submit(pinned host frames, pinned host labels): H2D / CUDA-graph replay / D2H on three "
                   "streams, 2 slots in flight; L2 flush inside the timed region")
        barrier()
        # the same pipeline fed DECODED frames (uint8 BGR HWC, what the reference's LoadImageFromFile produces): resize /
        # normalise / layout change on the GPU (vss_cffm_b200/preprocess.py), a quarter of the H2D bytes.  Reported beside
        # the headline e2e (whose input is the fp32 tensor the reference model itself receives).
        from vss_cffm_b200.preprocess import ClipPreprocessor
        pre = ClipPreprocessor()
        # ... and 8-bit label maps back (124 classes): an eighth of the D2H bytes; the values are the int64 labels' (checked below)
        pipe = ClipPipeline(model, B, T, H, W, metas, head_kw=dict(head_kw, label_dtype=torch.uint8), preprocessor=pre, src_hw=(H, W))
        g8 = torch.Generator().manual_seed(300 + rank)
        hosts = [torch.randint(0, 256, (T, B, H, W, 3), dtype=torch.uint8, generator=g8).pin_memory() for _ in range(3)]
        labs = [torch.empty(B, H, W, dtype=torch.uint8).pin_memory() for _ in range(3)]
        pipelined(3)
        fr = pre.run(hosts[2].cuda().view(T * B, H, W, 3), T, B)
        chk = model.labels_from_frames(fr, metas, **head_kw)
        torch.cuda.synchronize()
        check("uint8_pipeline_labels_equal_int64_labels", torch.equal(chk.cpu().to(torch.uint8), labs[2]))
        u8_ms = pipelined(args.steps)
        e2e_u8 = {"value": None, "unit": UNIT, "ms_per_step": round(u8_ms / args.steps, 4), "h2d_bytes_per_step": B * T * 3 * H * W,
                  "d2h_bytes_per_step": B * H * W,
                  "api": "ClipPipeline(preprocessor=ClipPreprocessor(), head_kw={'label_dtype': torch.uint8}).submit(pinned uint8 BGR HWC "
                         "frames, pinned uint8 labels)"}
        e2e_u8_ms = u8_ms
        barrier()

    # ---- streaming with temporal re-use (vss_cffm_b200/streaming.py; not the headline: the clips of the headline metric
    # are independent).  One step = the next frame of each of B videos; labels are bit-identical to the stateless path.
    # ---- the reference-facing call itself: model(img=[[T x (B,3,H,W)]], img_metas=[[...]], return_loss=False) -> list of numpy
    # label maps.  From the second call on it replays a cached CUDA graph (segmentor._cached_graph); every call still copies
    # its frames from host memory and returns host arrays, one call at a time (no overlap between calls).
    mmseg_call = None
    if not frames_mode and KIND == "cffm":
        for _ in range(3):
            model(img=[imgs_host], img_metas=[metas], return_loss=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            model(img=[imgs_host], img_metas=[metas], return_loss=False)
        mm_ms = (time.perf_counter() - t0) * 1e3
        mmseg_call = {"value": round(B * T * args.steps / (mm_ms * 1e-3), 2), "unit": UNIT, "ms_per_call": round(mm_ms / args.steps, 4),
                      "api": "EncoderDecoder_clips.forward(img, img_metas, return_loss=False) -> list[np.ndarray]; cached CUDA graph, "
                             "host wall clock, no L2 flush, one call at a time",
                      "cached_graphs": len(getattr(model, "_graphs", {}))}
    barrier()
    streaming = None
    if not frames_mode and T == 4 and KIND == "cffm":
        from vss_cffm_b200.streaming import VideoStream
        vs = VideoStream(model, B, graph=True)
        vframes = [imgs_dev[t % T] for t in range(8)]
        for t in range(2 * vs.history + 2):                      # fill the history and capture the per-residue graphs
            vs.push(vframes[t % 8])
        torch.cuda.synchronize()
        sm = timed(lambda: vs.push(vframes[vs.i % 8]), args.steps)
        streaming = {"value": round(B * args.steps / (sm * 1e-3), 2), "unit": "target-frames/s", "ms_per_step": round(sm / args.steps, 4),
                     "streams": B, "stateless_equivalent": None,
                     "note": "VideoStream.push: every frame is encoded once and its reference K/V cached; the stateless path "
                             "segments one target per T=4 clip-frames"}
    barrier()

    # per-kernel CUDA-event timing of the same step, launched eagerly (events cannot bracket nodes of a graph).  A spin
    # kernel is queued first so that the whole step (launches + events) is enqueued while the GPU is still busy: the
    # events then measure kernel durations, not the host's launch gaps.
    ksteps = min(args.steps, 5)
    with ops.KernelTimer() as kt:
        for _ in range(ksteps):
            flush.fill_(1)
            torch.cuda._sleep(int(12e-3 * 1.9e9))
            step_eager()
            torch.cuda.synchronize()
    krec = kt.results()
    eager_ms = sum(ms for _, _, ms in krec)                      # sum of the kernel durations of ksteps steps
    fam = {}
    for name, a, ms in krec:
        e = fam.setdefault(name.replace("cffm_", ""), [0, 0.0])
        e[0] += 1; e[1] += ms
    breakdown = {k: {"launches_per_step": v[0] // ksteps, "us_per_step": round(1e3 * v[1] / ksteps, 1)}
                 for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])}

    # ---- N > 1: the frame-sharded split of BASELINE configs[2] next to the clip-sharded headline: the frames of the SAME
    # global batch spread over the ranks, one NCCL all-gather of the reference-frame K/V per step (vss_cffm_b200/parallel.py)
    frame_shard = None
    if world > 1 and not frames_mode and T == 4 and KIND == "cffm":
        frame_shard = measure_frame_shard(args, model, rank, world, B, flush, serial_ms)   # like with like: one pass at a time
    t = torch.tensor([total_ms, e2e_ms, e2e_serial_ms, e2e_u8_ms, serial_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_serial_ms, e2e_u8_ms, serial_ms = t.tolist()
    frames = B * T * n_gpus * args.steps
    value = frames / (total_ms * 1e-3)
    e2e_value = frames / (e2e_ms * 1e-3)

    if rank == 0:
        pk = peaks()
        # DRAM bytes per launch measured by ncu for the same workload (profiles/r02_traffic.json, written by tools/ncu_traffic.py
        # from a committed capture; null when the file is missing)
        traffic = None
        for tf in ("r02_traffic.json", "r01_traffic.json"):
            try:
                traffic = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", tf)))
                break
            except Exception:
                continue

        def ncu_traffic(family):
            f = (traffic or {}).get("families", {}).get(family)
            return None if f is None else {"dram_bytes_per_launch": f["dram_bytes_per_launch"], "launches_per_step": f["launches_per_step"],
                                           "capture": traffic["source"]}
        # ---- dominant kernel family: the tcgen05 GEMM (every Linear / conv of the path).  Algorithmic bytes of a
        # launch = A + W + bias + residual + outputs, each once; FLOPs = 2 M N K.
        gem = {}
        g_bytes = g_flops = g_ms = 0.0
        for name, a, ms in krec:
            if name != "cffm_gemm_f16":
                continue
            Mg, Ng, Kg = a[11], a[12], a[13]
            by = 2 * Mg * Kg + 2 * Ng * Kg + (4 * Ng if a[4] else 0) + (4 * Mg * Ng if a[5] else 0) + \
                (2 * Mg * Ng if a[7] else 0) + (4 * Mg * Ng if a[9] else 0)
            fl = 2 * Mg * Ng * Kg
            g_bytes += by; g_flops += fl; g_ms += ms
            e = gem.setdefault((Mg, Ng, Kg), [0, 0.0, by, fl])
            e[0] += 1; e[1] += ms
        n_gemm = sum(e[0] for e in gem.values())
        top = sorted(gem.items(), key=lambda kv: -kv[1][1])[:5]
        gbs = g_bytes / (g_ms * 1e-3) / 1e9 if g_ms else 0.0
        roofline = {"kernel": "gemm_tcgen05_kernel (all launches of the step)", "bound": "hbm", "achieved": round(gbs, 1),
                    "peak": pk["hbm"], "unit": "GB/s", "frac": round(gbs / pk["hbm"], 4), "traffic": ncu_traffic("gemm_tcgen05_kernel"),
                    "tensor_tflops": round(g_flops / (g_ms * 1e-3) / 1e12, 2) if g_ms else 0.0,
                    "launches_per_step": n_gemm // ksteps, "ms_per_step": round(g_ms / ksteps, 4),
                    "share_of_kernel_time": round(g_ms / eager_ms, 4), "peak_source": pk["src"],
                    "algorithmic_bytes_per_step": int(g_bytes / ksteps), "algorithmic_flops_per_step": int(g_flops / ksteps),
                    "top_shapes": [{"M": k[0], "N": k[1], "K": k[2], "launches_per_step": v[0] // ksteps,
                                    "avg_us": round(1e3 * v[1] / v[0], 2), "GBps": round(v[2] / (v[1] / v[0] * 1e-3) / 1e9, 1),
                                    "frac": round(v[2] / (v[1] / v[0] * 1e-3) / 1e9 / pk["hbm"], 4)} for k, v in top],
                    "note": "K <= 256 for nearly every GEMM of the path (AI 30-120 FLOP/B < ridge 210): HBM is the roof. "
                            "Timed live with CUDA events around each launch of an eagerly launched step whose launches are queued behind a spin kernel (no host launch gaps inside the events)."}
        cfm_ms = [ms for name, a, ms in krec if name in ("cffm_cfm_attention", "cffm_cfm_attention_slots")]
        cfm_avg_ms = sum(cfm_ms) / max(len(cfm_ms), 1) if cfm_ms else float("nan")
        alg_bytes = CFM_BYTES_PER_CLIP_BLOCK * B                              # one launch = B clips of one block
        alg_flops = CFM_FLOPS_PER_CLIP_BLOCK * B
        cgbs = alg_bytes / (cfm_avg_ms * 1e-3) / 1e9
        ctfs = alg_flops / (cfm_avg_ms * 1e-3) / 1e12
        roofline_cfm = None if not cfm_ms else {"kernel": "cfm_attention_tc_kernel", "bound": "hbm", "achieved": round(cgbs, 2), "peak": pk["hbm"],
                        "unit": "GB/s", "frac": round(cgbs / pk["hbm"], 5), "traffic": ncu_traffic("cfm_attention_tc_kernel"),
                        "tensor_tflops": round(ctfs, 3), "tensor_frac_of_sustained": round(ctfs / pk["tf_sust"], 5),
                        "launch_ms": round(cfm_avg_ms, 5), "launches_timed": len(cfm_ms),
                        "share_of_kernel_time": round(sum(cfm_ms) / eager_ms, 4),
                        "algorithmic": {"bytes_per_launch": alg_bytes, "flops_per_launch": alg_flops,
                                        "note": "SURVEY.md 8(d): 9.6 MB and 1.1746 GFLOP per clip per block; un-fused attention "
                                                "has AI 122 FLOP/B < ridge 210, so HBM is the binding roof"}}
        # e2e headline: the pipeline fed DECODED frames (uint8, what the reference's loader hands to its transforms) when that
        # path ran; the same pipeline fed the normalised fp32 tensors the reference MODEL receives is reported beside it
        e2e_fp32 = {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(e2e_ms / args.steps, 4),
                    "h2d_bytes_per_step": B * T * 3 * H * W * 4, "d2h_bytes_per_step": B * H * W * 8, "api": e2e_api,
                    "input": "normalised fp32 frames (T x (B,3,H,W)), pinned host memory",
                    "serial_value": round(frames / (e2e_serial_ms * 1e-3), 2),
                    "serial_api": "graph.load(pinned host frames) -> replay -> D2H labels, one step at a time (no overlap)"}
        e2e_main = e2e_fp32
        if e2e_u8:
            e2e_main = dict(e2e_u8, value=round(frames / (e2e_u8_ms * 1e-3), 2), ms_per_step=round(e2e_u8_ms / args.steps, 4),
                            input="decoded uint8 BGR HWC frames (T,B,H,W,3), pinned host memory; AlignedResize_clips / Normalize_clips / "
                                  "ImageToTensor_clips run on the GPU inside the captured pass (bit-exact with cv2 / mmcv)",
                            serial_value=e2e_fp32["serial_value"], serial_api=e2e_fp32["serial_api"] + " (fp32 tensors)")
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": f"MiT-{VARIANT.upper()} + CFFM{'++ (K=%d prototypes)' % PROTOS if KIND == 'cffmpp' else ''} head (depth {HEAD_DEPTH[VARIANT]}), {H}x{W}, T={T}, {B} clips per GPU" +
                                   (" (BASELINE configs[1])" if (VARIANT, KIND, T, B) == ("b1", "cffm", 4, 2) else
                                    " (BASELINE configs[0]: T != num_clips, the head's early-return path)" if (VARIANT, T, B) == ("b0", 2, 1) else
                                    " (BASELINE configs[3])" if (VARIANT, KIND, T) == ("b2", "cffm", 4) else
                                    " (BASELINE configs[4])" if (VARIANT, KIND, PROTOS) == ("b1", "cffmpp", 64) else ""),
                       "clips_per_gpu": B, "frames_per_step": B * T * n_gpus, "shard": args.shard if n_gpus > 1 else "none",
                       "l2": "256 MiB flush per step" + (", inside the timed region (on the step's stream)" if passes_in_flight == 2 else ", between the per-step events"),
                       "timing": ("K steps as 2 passes in flight (one CUDA-graph replay each, own buffers, own stream), one CUDA-event pair around "
                                  "all K, max over ranks" if passes_in_flight == 2 else "per-step CUDA events, max over ranks"),
                       "passes_in_flight": passes_in_flight},
            "single_pass": {"value": round(frames / (serial_ms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(serial_ms / args.steps, 4),
                            "timing": "one pass at a time: per-step CUDA events, L2 flush between them (the round-1 definition of `value`)"},
            "clocks": clocks,
            "e2e": e2e_main,
            "e2e_from_fp32_tensors": e2e_fp32 if e2e_u8 else None,
            "e2e_mmseg_call": mmseg_call,
            "streaming": (dict(streaming, stateless_equivalent=round(value / T, 2)) if streaming else None),
            "gpu_launches": launches,
            "checks": checks,
            "launch_mode": (f"CUDA graph replay ({graphed.kernels_per_replay} kernel nodes per step)" if graphed is not None else
                            f"CUDA graph replay ({frames_graph_nodes} kernel nodes + 1 NCCL all-gather per step)" if frames_graph_nodes
                            else "eager"),
            "kernel_time_sum_ms_per_step": round(eager_ms / ksteps, 4),
            "kernel_breakdown": breakdown,
            "roofline": roofline,
            "roofline_cfm_attention": roofline_cfm,
        }
        if frame_shard is not None:
            out["frame_shard"] = frame_shard
        if not args.no_cpu_baseline and n_gpus == 1:              # reported on rank 0 at N = 1 only (the reference arm covers every N)
            v, ms, cores = cpu_reference_time(sd, 3, 1, clips=1)
            out["cpu_baseline"] = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"1 clip (T={T}, {H}x{W}), 3 timed forwards after 1 warm-up, fp32 CPU oracle, "
                                             f"{ms:.0f} ms per clip"}
        print(json.dumps(out))
    if world > 1:
        if gfs is not None:
            gfs.close()                                          # a graph with a captured collective must go before the communicator
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
