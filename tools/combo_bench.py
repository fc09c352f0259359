#!/usr/bin/env python
"""COMBO-R50 S4 model-level measurements with the product's pixel decoder on the hot path (harness/: SURVEY.md 8 f3 / f4).

  python tools/combo_bench.py train [clips_per_gpu=8] [iters=8] [strong] [pvt] [ms3|avss]   # GPU(s): full training step, DDP under torchrun
  python tools/combo_bench.py cpu-forward [clips=1]                        # BASELINE config 1: eval forward on the host CPU

train: forward + matcher + criterion + backward + full-model gradient clipping (0.01) + AdamW (train_net.py:147-226), fp32 (AMP is
off for S4: configs/avs_s4/R50-AVSS4-SemanticSegmentation.yaml:44-45), 5-frame 224x224 synthetic clips, first frame supervised.
`pvt` swaps both backbones for PVTv2-B5 (harness/pvtv2.py); `ms3` supervises all five frames (BASELINE config 4); `avss` is config 5:
71 classes, 10-frame clips, every frame supervised, fp16 autocast + GradScaler around the step (AMP is on for AVSS:
configs/avs_ss/PVT-AVSS-SemanticSegmentation.yaml:41-42) while the pixel decoder keeps float32 (msdeformattn.py:315,321).
Under torchrun every rank runs its own clips (weak) or, with `strong`, its shard of a global batch of 8 clips (train_net.py:117).
cpu-forward: the same model on the CPU with the hot path served by the reference's CPU implementation (ms_deform_attn_core_pytorch,
restated in oracle/core_pytorch.py) -- a BASELINE measurement, which is why this tool (and nothing in the product) touches oracle/."""
import json
import os
import sys
import time

os.environ["NCCL_DEBUG_FILE"] = os.environ.get("NCCL_DEBUG_FILE", "/dev/stderr")
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def train(clips, iters, strong, backbone="r50", task="s4"):
    import combo_avs_b200 as pkg
    from combo_avs_b200.sharding import max_over_ranks, shard_clips
    from harness.combo_model import ComboR50, build_optimizer, clip_full_model_gradients, synthetic_clips
    from harness.criterion import SetCriterion
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    frames_per_clip = 10 if task == "avss" else 5
    num_classes = 71 if task == "avss" else 2
    supervised = 1 if task == "s4" else frames_per_clip
    amp = task == "avss"
    if strong:
        clips = shard_clips(8, frames_per_clip, rank, world)[1] // frames_per_clip
    torch.manual_seed(7)
    model = ComboR50(num_classes=num_classes, num_frames=frames_per_clip, backbone=backbone).to(dev).train()
    criterion = SetCriterion(num_classes=num_classes, num_frames=frames_per_clip).to(dev)
    scaler = torch.amp.GradScaler("cuda", enabled=amp)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = build_optimizer(model)
    images, pre_masks, audio, targets = synthetic_clips(clips, num_frames=frames_per_clip, device=dev, seed=100 + rank,
                                                        supervised_frames=supervised, num_classes=num_classes)
    times = {}

    def step(record=False):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
        opt.zero_grad(set_to_none=True)
        if record: ev[0].record()
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            out = ddp(images, pre_masks, audio)
            if record: ev[1].record()
            loss = criterion(out, targets)
        if record: ev[2].record()
        scaler.scale(loss).backward()
        if record: ev[3].record()
        scaler.unscale_(opt)
        clip_full_model_gradients(model, 0.01)
        scaler.step(opt)
        scaler.update()
        if record:
            ev[4].record()
            torch.cuda.synchronize()
            for k, (a, b) in zip(("forward", "matcher+criterion", "backward(+allreduce)", "clip+adamw"), zip(ev[:-1], ev[1:])):
                times.setdefault(k, []).append(a.elapsed_time(b))
        return loss

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = pkg.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / iters
    launches = (pkg.launch_count() - l0) / iters
    for _ in range(3):
        step(record=True)
    frames = (8 * frames_per_clip if strong else world * clips * frames_per_clip)
    res = {"workload": "COMBO-%s %s training step (harness/ + product pixel decoder), %s, %d-frame 224x224 synthetic clips, %s supervised" % (
               "PVTv2-B5" if backbone == "pvtv2_b5" else "R50", task.upper(), "fp16 autocast (pixel decoder fp32)" if amp else "fp32", frames_per_clip,
               "first frame" if supervised == 1 else "every frame"),
           "n_gpus": world, "scaling": "strong (global batch 8 clips)" if strong else "weak (%d clips per GPU)" % clips,
           "clips_per_gpu": clips, "global_frames": frames, "ms_per_step": ms, "frames_per_s": frames / (ms * 1e-3),
           "loss": float(loss), "msda_launches_per_step": launches,
           "parameters": sum(p.numel() for p in model.parameters()), "trainable_parameters": sum(p.numel() for p in model.parameters() if p.requires_grad),
           "allreduce_bytes_per_step": sum(p.numel() * 4 for p in model.parameters() if p.requires_grad) if world > 1 else 0,
           "phases_ms_rank0": {k: sum(v) / len(v) for k, v in times.items()}}
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_forward(clips):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from core_pytorch import msda_core_pytorch
    from combo_avs_b200 import modules as pkg_modules
    from harness.combo_model import ComboR50, synthetic_clips

    class _CpuReferenceFunction:  # the reference module's CPU route (ms_deform_attn.py:123-125 -> ms_deform_attn_core_pytorch)
        @staticmethod
        def apply(value, shapes, start, loc, attn, step):
            t0 = time.perf_counter()
            out = msda_core_pytorch(value, shapes, loc, attn)
            _CpuReferenceFunction.seconds += time.perf_counter() - t0
            return out
        seconds = 0.0

    pkg_modules.MSDeformAttnFunction = _CpuReferenceFunction
    torch.manual_seed(7)
    model = ComboR50().eval()
    for layer in model.pixel_decoder.transformer.encoder.layers:
        layer.self_attn.fuse_prologue = False
    images, pre_masks, audio, _ = synthetic_clips(clips, device="cpu", seed=100)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with torch.no_grad():
        model(images, pre_masks, audio)  # warm-up
        _CpuReferenceFunction.seconds = 0.0
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            model(images, pre_masks, audio)
        sec = (time.perf_counter() - t0) / reps
    print(json.dumps({"workload": "COMBO-R50 S4 eval forward on CPU (BASELINE config 1), hot path = ms_deform_attn_core_pytorch (oracle/core_pytorch.py)",
                      "clips": clips, "frames": clips * 5, "cores": threads, "torch": torch.__version__, "seconds_per_clip": sec / clips,
                      "msdeformattn_core_share": _CpuReferenceFunction.seconds / reps / sec}), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "train"
    if mode == "cpu-forward":
        cpu_forward(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    else:
        flags = sys.argv[4:]
        train(int(sys.argv[2]) if len(sys.argv) > 2 else 8, int(sys.argv[3]) if len(sys.argv) > 3 else 8, "strong" in flags,
              backbone="pvtv2_b5" if "pvt" in flags else "r50", task="avss" if "avss" in flags else "ms3" if "ms3" in flags else "s4")
