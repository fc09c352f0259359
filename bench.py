#!/usr/bin/env python
"""bench.py -- MSDeformAttn core fwd+bwd throughput on B200 (BASELINE.json configs[1]), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype fp32|bf16] [--frames F]

A "step" is one forward + one backward of the MSDeformAttn core over one batch of synthetic S4-shaped frames
(N=40 frames = 8 clips x 5 frames, 3 levels 7x7/14x14/28x28, 8 heads, 4 points, d_model 256, Lq = S = 1029),
inputs drawn like the reference's own test (ops/test.py:36-39).  `value` is frames/s through fwd+bwd with inputs
resident in HBM; `e2e` is the same metric through the host-buffer entry point (pinned host tensors -> H2D -> kernels
-> D2H of every result, all inside the timed region).  N>1: one process per GPU (torchrun), every rank runs its own
40-frame batch (weak scaling, the op has no cross-frame dependency and no collective); time = max over ranks.

--impl reference times the reference's CPU implementation of the same path (ms_deform_attn_core_pytorch,
restated in oracle/core_pytorch.py because /root/reference does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: NCCL's banner ("NCCL version ..." under NCCL_DEBUG=VERSION/INFO) is written to stdout by
# the library itself, so its debug output is pointed at stderr before torch / NCCL are loaded
if (int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN")
        and "NCCL_DEBUG_FILE" not in os.environ):
    # multi-GPU run: ask NCCL's tuner to log which all-reduce algorithm it picks (per-rank file, parsed by nccl_algorithm_seen)
    os.environ["MSDA_NCCL_LOG"] = "/tmp/msda_bench_nccl_%h_%p.log"
    os.environ["NCCL_DEBUG_FILE"] = os.environ["MSDA_NCCL_LOG"]
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,TUNING"
os.environ["NCCL_DEBUG_FILE"] = os.environ.get("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S4_SHAPES = [(7, 7), (14, 14), (28, 28)]
M, D, L, P = 8, 32, 3, 4
S = sum(h * w for h, w in S4_SHAPES)
LQ = S
METRIC = "msdeformattn_fwd_bwd_frames_per_s"
UNIT = "frames/s"
N_SETS = 4  # rotating input/output sets so that every step starts cold in L2


def algorithmic_bytes(n_frames, val_bytes, meta_bytes):
    """every input read once + every output written once, per pass."""
    value = n_frames * S * M * D * val_bytes
    loc = n_frames * LQ * M * L * P * 2 * meta_bytes
    attn = n_frames * LQ * M * L * P * meta_bytes
    out = n_frames * LQ * M * D * val_bytes
    fwd = value + loc + attn + out
    bwd = (value + loc + attn) + out + value + (loc + attn)  # inputs + grad_out + grad_value + grad_loc + grad_attn
    return fwd, bwd


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(variant):
    """DRAM bytes per launch of a kernel variant from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(variant)
    except Exception:
        return None


def synth_inputs(torch, n_frames, seed, dtype, loc_dtype, device, pin=False):
    g = torch.Generator().manual_seed(seed)
    value = torch.rand(n_frames, S, M, D, generator=g) * 0.01
    loc = torch.rand(n_frames, LQ, M, L, P, 2, generator=g)
    attn = torch.rand(n_frames, LQ, M, L, P, generator=g) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    gout = torch.rand(n_frames, LQ, M * D, generator=g)
    ts = [value.to(dtype), loc.to(loc_dtype), attn.to(loc_dtype), gout.to(dtype)]
    if device is not None:
        return [t.to(device) for t in ts]
    return [t.pin_memory() if pin else t for t in ts]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run (B200_PROFILING.md 'clocks line')."""
    FIELDS = ("uuid,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid = uuid
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9 and (self.uuid is None or self.uuid in parts[0]):
                self.rows.append((time.time(), parts))

    def stop(self, windows):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [p for t, p in self.rows if any(a <= t <= b for a, b in windows)] or [p for _, p in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(p[1]) for p in rows if num(p[1]) is not None]
        reasons = set()
        for p in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": num(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max((num(p[3]) or 0.0) for p in rows)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def time_cpu_port(torch, n_frames, steps, warmup, threads=None):
    """fwd+bwd of the reference's CPU path (restated: oracle/core_pytorch.py <- ms_deform_attn_func.py:53-72) in fp32."""
    from oracle.core_pytorch import msda_core_pytorch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    value, loc, attn, gout = synth_inputs(torch, n_frames, 3, torch.float32, torch.float32, None)
    shapes = torch.tensor(S4_SHAPES, dtype=torch.long)
    value.requires_grad_(True)
    loc.requires_grad_(True)
    attn.requires_grad_(True)

    def step():
        value.grad = loc.grad = attn.grad = None
        out = msda_core_pytorch(value, shapes, loc, attn)
        out.backward(gout)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dt / steps, threads


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_frames = args.frames
    # bounded sample: shrink the batch until one fwd+bwd takes <= ~1.5 s on this host
    probe, threads = time_cpu_port(torch, min(n_frames, 5), 1, 1)
    per_frame = probe / min(n_frames, 5)
    sample = n_frames
    while sample > 5 and per_frame * sample > 1.5:
        sample //= 2
    sec, threads = time_cpu_port(torch, sample, args.steps, args.warmup)
    fps = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_frames, "f32"),
        "note": "one host CPU complex times a bounded sample of ONE rank's batch whatever --gpus says (the CPU path does not shard "
                "over GPUs): compare it with the N=1 line; at N>1 the ratio is N GPUs against the same single host",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d of %d frames per step, fwd+bwd, torch %s CPU grid_sample path (oracle/core_pytorch.py)" % (sample, n_frames, torch.__version__)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_frames, dtype):
    return {"workload": "standalone MSDeformAttn core fwd+bwd, AVSBench S4 shape (BASELINE.json configs[1])",
            "frames_per_gpu": n_frames, "levels": S4_SHAPES, "heads": M, "points": P, "head_dim": D, "queries": LQ,
            "im2col_step": 128, "inputs": "value=U*0.01, loc=U[0,1), attn=normalised U (reference ops/test.py:36-39), seed 3",
            "l2": "%d rotating input/output sets (> 126 MB L2 each step cold)" % N_SETS, "value_dtype": dtype}



# ------------------------------------------------------------------------------------- the reference's CUDA kernels, same tensors
def time_incumbent(torch, pkg, sets, shapes, start, n_frames, dtype, loc_dtype, windows):
    """fwd / bwd microseconds of the reference's own extension (ops/src/cuda/ms_deform_im2col_cuda.cuh:242-408, built unmodified
    but for a two-line torch-2.x patch by oracle/build_ref.py into oracle/_ref/) and its max-norm relative difference from ours."""
    if dtype != torch.float32 or loc_dtype != torch.float32:
        return {"unavailable": "the reference extension is float32 / float64 only"}
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import build_ref
        ref = build_ref.load_ref()
    except Exception as e:  # pragma: no cover - depends on the box
        return {"unavailable": "oracle/_ref failed to load: %s" % e}
    finally:
        sys.path.pop(0)
    if ref is None:
        return {"unavailable": "oracle/_ref/MultiScaleDeformableAttention.so did not travel (build it with python oracle/build_ref.py)"}

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

    v, l, a, g = sets[0]
    out_r = ref.ms_deform_attn_forward(v, shapes, start, l, a, 128)
    gv_r, gl_r, ga_r = ref.ms_deform_attn_backward(v, shapes, start, l, a, g, 128)
    out = pkg.ms_deform_attn_forward(v, shapes, start, l, a, 128)
    gv, gl, ga = pkg.ms_deform_attn_backward(v, shapes, start, l, a, g, 128)
    diffs = {"out": rel(out, out_r), "grad_value": rel(gv, gv_r), "grad_sampling_loc": rel(gl, gl_r), "grad_attn_weight": rel(ga, ga_r)}
    del out_r, gv_r, gl_r, ga_r, out, gv, gl, ga

    def timed(fn, reps):
        for i in range(3):
            fn(*sets[i % N_SETS])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(reps):
            fn(*sets[i % N_SETS])
        e1.record()
        torch.cuda.synchronize()
        windows.append((t0, time.time()))
        return e0.elapsed_time(e1) / reps * 1e3

    us_f = timed(lambda v, l, a, g: ref.ms_deform_attn_forward(v, shapes, start, l, a, 128), 40)
    us_b = timed(lambda v, l, a, g: ref.ms_deform_attn_backward(v, shapes, start, l, a, g, 128), 20)
    return {"what": "the reference's own CUDA extension (oracle/_ref, ms_deform_im2col_cuda.cuh kernels) on the same tensors and GPU",
            "fwd_us": us_f, "bwd_us": us_b, "fwd_bwd_us": us_f + us_b, "frames_per_s": n_frames / ((us_f + us_b) * 1e-6),
            "max_rel_diff_vs_ours": diffs, "allocations": "callee-allocated outputs (at::zeros), as the reference does"}


# --------------------------------------------------------------------------- the caller: pixel-decoder DDP training step
def time_model_step(torch, dist, pkg, dev, rank, world, windows, iters=8):
    """Training step (forward + backward + fused AdamW, fp32 like the reference: AMP is off for S4/MS3 and msdeformattn.py:321 forces
    float) of the COMBO-R50 S4 MSDeformAttnPixelDecoder (msdeformattn.py:169-368: input projections + GroupNorm, 6 deformable encoder
    layers = 6 calls of the op each way, FPN top-down pass, mask features) on synthetic ResNet-50-shaped features of 224x224 frames,
    one process per GPU under DistributedDataParallel.  weak: 8 clips x 5 frames per GPU; strong: a global batch of 8 clips sharded
    by clip (detectron2 semantics, train_net.py:117).  The backbones, audio branch, transformer decoder and criterion of the full
    model are NOT in this step (tools/combo_harness.py holds the full-model harness)."""
    from combo_avs_b200.pixel_decoder import MSDeformAttnPixelDecoder
    from combo_avs_b200.sharding import max_over_ranks, shard_clips
    r50 = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}
    torch.manual_seed(7)
    top = MSDeformAttnPixelDecoder(r50, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                   transformer_enc_layers=6, conv_dim=256, mask_dim=256, norm="GN",
                                   transformer_in_features=["res3", "res4", "res5"], common_stride=4).to(dev)
    with torch.no_grad():  # off the zero init, so offsets / weights vary per query as in a trained model
        for layer in top.transformer.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
            layer.self_attn.attention_weights.weight.normal_(0, 0.1)
    model = torch.nn.parallel.DistributedDataParallel(top, device_ids=[dev.index]) if world > 1 else top
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
    grad_bytes = sum(p.numel() * p.element_size() for p in top.parameters() if p.requires_grad)
    res = {"model": "COMBO-R50 S4 MSDeformAttnPixelDecoder (6 deformable encoder layers, conv_dim 256, FFN 1024), fp32, fwd+bwd+AdamW",
           "unit": "frames/s", "n_gpus": world, "parameters": sum(p.numel() for p in top.parameters()),
           "allreduce_bytes_per_step": grad_bytes if world > 1 else 0,
           "collective": "none (1 GPU)"}
    l0 = pkg.launch_count()
    for mode in ("weak", "strong"):
        frames = 40 if mode == "weak" else shard_clips(8, 5, rank, world)[1]
        if mode == "strong" and world == 1:
            res[mode] = dict(res["weak"])
            continue
        g = torch.Generator().manual_seed(100 + rank)
        feats = {k: torch.randn(max(frames, 1), c, 224 // s, 224 // s, generator=g).to(dev) for k, (c, s) in r50.items()}

        def step():
            opt.zero_grad(set_to_none=True)
            mask_feat, low, multi = model(feats)
            (mask_feat.mean() + low.mean() + sum(m.mean() for m in multi)).backward()
            opt.step()

        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        windows.append((t0, time.time()))
        ms = max_over_ranks(e0.elapsed_time(e1), dev) / iters
        total_frames = world * 40 if mode == "weak" else 40
        res[mode] = {"frames_per_gpu": frames, "global_frames": total_frames, "ms_per_step": ms, "value": total_frames / (ms * 1e-3)}
        del feats
    res["msda_launches_per_step"] = (pkg.launch_count() - l0) / ((2 if world > 1 else 1) * (iters + 3))
    if world > 1:  # (asked after the all-reduces have run: NCCL's tuner logs its choice at the first collective of each size)
        res["collective"] = ("DistributedDataParallel gradient all-reduce, NCCL %s, bucket_cap 25 MiB; algorithm: %s"
                             % (".".join(map(str, torch.cuda.nccl.version())), nccl_algorithm_seen()))
    del model, top, opt
    torch.cuda.empty_cache()
    try:
        res["full_model"] = time_full_model_step(torch, dist, pkg, dev, rank, world, windows)
    except Exception as e:  # the harness is measurement scaffolding: its failure must not take the headline down
        res["full_model"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    torch.cuda.empty_cache()
    return res


def time_full_model_step(torch, dist, pkg, dev, rank, world, windows, iters=5):
    """BASELINE.json config 3: the COMBO-R50 S4 training step -- two ResNet-50 trunks + channel gates, frozen VGGish, the product's
    pixel decoder (the hot path), late AVFuse, 9-layer masked transformer decoder, batched Hungarian matcher + criterion, full-model
    gradient clipping, AdamW -- on synthetic 5-frame 224x224 clips with the first frame supervised, fp32 (AMP is off for S4), DDP.
    The model around the pixel decoder is the detectron2-free restatement in harness/ (random init); frames/s counts all 5 frames."""
    from combo_avs_b200.sharding import max_over_ranks, shard_clips
    from harness.combo_model import ComboR50, build_optimizer, clip_full_model_gradients, synthetic_clips
    from harness.criterion import SetCriterion
    torch.manual_seed(7)
    model = ComboR50().to(dev).train()
    criterion = SetCriterion().to(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if world > 1 else model
    opt = build_optimizer(model)
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    res = {"model": "COMBO-R50 S4 (harness/combo_model.py + product pixel decoder), fp32, fwd + matcher/criterion + bwd + clip 0.01 + AdamW",
           "unit": "frames/s", "n_gpus": world, "trainable_parameters": trainable, "allreduce_bytes_per_step": trainable * 4 if world > 1 else 0}
    for mode in ("weak", "strong"):
        clips = 8 if mode == "weak" else shard_clips(8, 5, rank, world)[1] // 5
        if mode == "strong" and world == 1:
            res[mode] = dict(res["weak"])
            continue
        images, pre_masks, audio, targets = synthetic_clips(clips, device=dev, seed=100 + rank)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = criterion(ddp(images, pre_masks, audio), targets)
            loss.backward()
            clip_full_model_gradients(model, 0.01)
            opt.step()

        for _ in range(2):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        windows.append((t0, time.time()))
        ms = max_over_ranks(e0.elapsed_time(e1), dev) / iters
        total = world * 40 if mode == "weak" else 40
        res[mode] = {"clips_per_gpu": clips, "global_frames": total, "ms_per_step": ms, "value": total / (ms * 1e-3)}
        del images, pre_masks, audio, targets
    del ddp, model, opt, criterion
    return res


def nccl_algorithm_seen():
    """Best effort: the algorithm NCCL's tuner picked for the largest all-reduce, parsed from this rank's NCCL_DEBUG=INFO /
    NCCL_DEBUG_SUBSYS=TUNING log when bench.py was allowed to request one (see main)."""
    path = os.environ.get("MSDA_NCCL_LOG")
    names = {"0": "Tree", "1": "Ring", "2": "CollnetDirect", "3": "CollnetChain", "4": "NVLS", "5": "NVLSTree", "6": "PAT"}
    protos = {"0": "LL", "1": "LL128", "2": "Simple"}
    best = None
    try:
        import glob
        import re
        for f in glob.glob(path.replace("%h", "*").replace("%p", "*")) if path else []:
            for line in open(f, errors="replace"):
                m = re.search(r"AllReduce: (\d+) Bytes -> Algo (\w+) proto (\w+)", line)
                if m and (best is None or int(m.group(1)) > best[0]):
                    best = (int(m.group(1)), m.group(2), m.group(3))
    except Exception:
        best = None
    if best is None:
        return "chosen by NCCL's tuner (not logged in this run)"
    return "%s, protocol %s for the %d-byte bucket (NCCL tuner log)" % (names.get(best[1], best[1]), protos.get(best[2], best[2]), best[0])


# ------------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:  # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import combo_avs_b200 as pkg
    from combo_avs_b200.host_api import HostPipeline
    from combo_avs_b200.sharding import max_over_ranks

    if args.fwd_variant:
        pkg.select_variant("forward", args.fwd_variant)
    if args.bwd_variant:
        pkg.select_variant("backward", args.bwd_variant)
    dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
    loc_dtype = torch.float32 if args.dtype == "fp32" or args.loc_dtype == "fp32" else torch.bfloat16
    nF = args.frames
    shapes = torch.tensor(S4_SHAPES, dtype=torch.long, device=dev)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    sets = [synth_inputs(torch, nF, 3 + i + 100 * rank, dtype, loc_dtype, dev) for i in range(N_SETS)]
    fwd = pkg.ms_deform_attn_forward
    bwd = pkg.ms_deform_attn_backward

    def step(i):
        v, l, a, g = sets[i % N_SETS]
        out = fwd(v, shapes, start, l, a, 128)
        return out, bwd(v, shapes, start, l, a, g, 128)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1), dev)
        return ms, (t0, time.time())

    uuid = None
    try:
        uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        pass
    sampler = ClockSampler(uuid) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.5)  # let nvidia-smi finish initialising (it holds driver locks while it does) before anything is timed
    windows = []

    # CUDA-graph replay of a launch sequence: the N_SETS launches of `one(i)` (one per rotating input set) are captured once and replayed;
    # a step count that is not a multiple of N_SETS ends with eager launches.  Returns (run_k, launches_per_call_of_one) or None.
    def replayer(one):
        try:
            one(0)
            torch.cuda.synchronize()
            cg = torch.cuda.CUDAGraph()
            keep = []
            lc = pkg.launch_count()
            with torch.cuda.graph(cg):
                for i in range(N_SETS):
                    keep.append(one(i))
            per_call = (pkg.launch_count() - lc) / N_SETS

            def run_k(k):
                for _ in range(k // N_SETS):
                    cg.replay()
                for i in range(k % N_SETS):
                    one(i)
            run_k.keep = (cg, keep)
            return run_k, per_call
        except Exception as e:  # capture refused: the eager loop below is the fallback, and the line says so
            sys.stderr.write("bench.py: CUDA-graph capture failed (%s: %s); timing eager launches\n" % (type(e).__name__, e))
            torch.cuda.synchronize()
            return None

    def timed_k(run_k, k, warmup):
        run_k(max(warmup, N_SETS) + k % N_SETS)  # (the eager remainder of the timed call, if any, is warmed too)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        run_k(k)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev), (t0, time.time())

    def fwd_only(i):
        return fwd(sets[i % N_SETS][0], shapes, start, sets[i % N_SETS][1], sets[i % N_SETS][2], 128)

    def bwd_only(i):
        return bwd(sets[i % N_SETS][0], shapes, start, sets[i % N_SETS][1], sets[i % N_SETS][2], sets[i % N_SETS][3], 128)

    # ---- the headline: K fwd+bwd steps, inputs resident in HBM ------------------------------------------------------
    # Launched the way a training loop that cares launches them -- replayed from a CUDA graph (the C ABI allocates nothing and never
    # synchronises, so the op records as is; tests/test_gpu_parity.py::test_cuda_graph_capture_and_replay) -- because a 5-50 ms timed
    # region of eager launches is at the mercy of the host (Python call + three cuTensorMapEncodeTiled per launch, an nvidia-smi query
    # taking the driver lock: the same kernels gave 244 and 259 us per step in two runs).  The eager loop is reported next to it.
    warm = max(args.warmup, 3)
    l0 = pkg.launch_count()
    ms_eager, w = timed(step, args.steps, warm)
    launches_eager = (pkg.launch_count() - l0) * args.steps // (args.steps + warm)
    windows.append(w)
    eager = {"ms_per_step": ms_eager / args.steps, "value": world * nF / (ms_eager / args.steps * 1e-3), "unit": UNIT, "gpu_launches": int(launches_eager),
             "what": "the same K steps as eager launches from Python (one ctypes call per kernel)"}
    rep_step = None if args.eager else replayer(step)
    if rep_step is not None:
        ms_total, w = timed_k(rep_step[0], args.steps, warm)
        windows.append(w)
        launches = int(round(rep_step[1] * args.steps))
        launch_mode = "cuda_graph_replay: %d steps (one per rotating input set) captured once with torch.cuda.graph, replayed %d times%s" % (
            N_SETS, args.steps // N_SETS, "" if args.steps % N_SETS == 0 else " + %d eager steps" % (args.steps % N_SETS))
    else:
        ms_total, launches, launch_mode = ms_eager, launches_eager, "eager launches from Python"
    ms_step = ms_total / args.steps
    fps = world * nF / (ms_step * 1e-3)

    # ---- per-kernel durations for the roofline (rank 0's GPU, same rotating sets, same launch mode) ------------------------
    reps = max(args.steps, int(0.4 / max(ms_step * 1e-3, 1e-6)))  # >= ~0.4 s per loop so the clock sampler sees it
    rep_f = None if rep_step is None else replayer(fwd_only)
    rep_b = None if rep_step is None else replayer(bwd_only)
    if rep_f is not None and rep_b is not None:
        ms_f, w = timed_k(rep_f[0], reps, 3)
        windows.append(w)
        ms_b, w = timed_k(rep_b[0], reps, 3)
        windows.append(w)
    else:
        ms_f, w = timed(fwd_only, reps, 3)
        windows.append(w)
        ms_b, w = timed(bwd_only, reps, 3)
        windows.append(w)
    us_f, us_b = ms_f / reps * 1e3, ms_b / reps * 1e3
    vb, mb = (4 if dtype == torch.float32 else 2), (4 if loc_dtype == torch.float32 else 2)
    bytes_f, bytes_b = algorithmic_bytes(nF, vb, mb)
    peak, peak_src = load_peaks()
    dom = ("backward", bytes_b, us_b) if us_b >= us_f else ("forward", bytes_f, us_f)
    achieved = dom[1] / (dom[2] * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": "%s (%s)" % (dom[0], pkg.last_variant(dom[0])), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": (load_traffic(pkg.last_variant(dom[0]) + ("" if args.dtype == "fp32" else "_bf16")) or {}).get("dram_bytes_per_launch")
                if nF == 40 and (args.dtype == "fp32" or loc_dtype == torch.float32) else None,
                "traffic_source": "constant: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture "
                                  "(profiles/traffic.json), not measured in this run",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom[1], "us_per_launch": dom[2],
                "fwd": {"us": us_f, "bytes": bytes_f, "GBps": bytes_f / us_f / 1e3, "frac": bytes_f / us_f / 1e3 / peak, "variant": pkg.last_variant("forward")},
                "bwd": {"us": us_b, "bytes": bytes_b, "GBps": bytes_b / us_b / 1e3, "frac": bytes_b / us_b / 1e3 / peak, "variant": pkg.last_variant("backward")},
                "fwd_bwd": {"us": us_f + us_b, "bytes": bytes_f + bytes_b, "frac": (bytes_f + bytes_b) / (us_f + us_b) / 1e3 / peak}}
    rep_step = rep_f = rep_b = None  # (drop the captured graphs and their private pools before the other legs allocate)

    # ---- end to end through the host-buffer API: pinned host inputs -> H2D -> kernels -> D2H of all results ------------
    host_in = synth_inputs(torch, nF, 3 + 100 * rank, dtype, loc_dtype, None, pin=True)
    host_out = [torch.empty((nF, LQ, M * D), dtype=dtype).pin_memory(), torch.empty_like(host_in[0]).pin_memory(),
                torch.empty_like(host_in[1]).pin_memory(), torch.empty_like(host_in[2]).pin_memory()]
    pipe = HostPipeline(dev, chunk_frames=args.e2e_chunk)
    xfer = {}

    def e2e_step(i):
        xfer["h2d"], xfer["d2h"] = pipe.forward_backward(host_in[0], shapes, start, host_in[1], host_in[2], host_in[3], *host_out, 128)

    e2e_steps = max(3, min(args.steps, 50))
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.time()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    t1 = time.time()
    e2e_s = (t1 - t0) / e2e_steps  # host wall clock: the call returns only when results are on the host
    e2e_s = max_over_ranks(e2e_s, dev)
    windows.append((t0, t1))
    e2e = {"value": world * nF / e2e_s, "unit": UNIT, "h2d_bytes_per_step": xfer["h2d"], "d2h_bytes_per_step": xfer["d2h"],
           "ms_per_step": e2e_s * 1e3, "api": "combo_avs_b200.host_api.HostPipeline.forward_backward -> msda_forward_backward_host (C ABI; pinned host tensors, %d-frame chunks, 3 streams, 3 staging slots)" % args.e2e_chunk}

    # ---- the incumbent: the reference's own CUDA kernels (oracle/_ref, compiled from /root/reference by oracle/build_ref.py) on the
    # same tensors, same GPU -- measurement infrastructure only, nothing of it is on the product path
    incumbent = None
    if rank == 0 and world == 1 and not args.no_incumbent:
        incumbent = time_incumbent(torch, pkg, sets, shapes, start, nF, dtype, loc_dtype, windows)

    # ---- one level up: the DDP training step of the pixel decoder that calls the op 6x per pass (weak: 8 clips per GPU; strong:
    # a global batch of 8 clips sharded by clip, train_net.py:117) -- NCCL gradient all-reduce is the only collective
    model_step = None
    if not args.no_model_step and dtype == torch.float32:
        model_step = time_model_step(torch, dist, pkg, dev, rank, world, windows)

    clocks = sampler.stop(windows) if sampler else None

    # ---- CPU baseline beside it (rank 0, single-GPU run only): bounded sample of the same workload ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = min(nF, 40)
        probe, threads = time_cpu_port(torch, 5, 1, 1)
        iters = max(2, min(10, int(15.0 / max(probe / 5 * sample, 1e-3))))
        sec, threads = time_cpu_port(torch, sample, iters, 1)
        cpu = {"value": sample / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d frames x %d fwd+bwd steps, fp32, torch %s CPU path restating ms_deform_attn_core_pytorch (oracle/core_pytorch.py)" % (sample, iters, torch.__version__),
               "ms_per_step": sec * 1e3}
        # the same path on ONE host thread as well (2-frame sample), and which CPU this is
        sec1, _ = time_cpu_port(torch, 2, 2, 1, threads=1)
        cpu["single_thread"] = {"value": 2 / sec1, "unit": UNIT, "cores": 1, "sample": "2 frames x 2 fwd+bwd steps"}
        torch.set_num_threads(threads)
        try:
            with open("/proc/cpuinfo") as f:
                cpu["cpu_model"] = next((l.split(":", 1)[1].strip() for l in f if l.startswith("model name")), None)
        except OSError:
            cpu["cpu_model"] = None

    if rank == 0:
        line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == torch.float32 else "bf16",
                "data": "synthetic", "config": workload_config(nF, "f32" if dtype == torch.float32 else "bf16"), "launch": launch_mode, "fwd_bwd_us": ms_step * 1e3,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "incumbent": incumbent, "model_step": model_step, "eager_launch_loop": eager,
                "variants": {"forward": pkg.last_variant("forward"), "backward": pkg.last_variant("backward")}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--loc-dtype", default="fp32", choices=["fp32", "bf16"], help="bf16 runs only: dtype of locations / weights")
    ap.add_argument("--frames", type=int, default=40, help="frames per GPU per step (S4: 8 clips x 5 frames)")
    ap.add_argument("--fwd-variant", default=None)
    ap.add_argument("--bwd-variant", default=None)
    ap.add_argument("--e2e-chunk", type=int, default=5)
    ap.add_argument("--eager", action="store_true", help="time eager launches only (no CUDA-graph replay of the step)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the reference-CUDA-extension leg")
    ap.add_argument("--no-model-step", action="store_true", help="skip the pixel-decoder DDP training-step leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
