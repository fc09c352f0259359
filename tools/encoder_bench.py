#!/usr/bin/env python
"""GPU: the caller of the hot path at its production shape -- the 6-layer deformable encoder of the pixel decoder
(msdeformattn.py:23-166; d_model 256, 8 heads, FFN 1024, 3 levels 7x7/14x14/28x28) -- training step on synthetic features:
forward + backward (+ AdamW) over N frames per GPU, fp32 like the reference (AMP is off for S4/MS3 and the pixel decoder
forces float anyway, msdeformattn.py:321).  Arms, same PyTorch code around the op:
    fused      -- this package, module prologue + mask inside the kernels (default)
    unfused    -- this package, softmax / locations / masked_fill in PyTorch
    reference  -- the reference's own CUDA extension (oracle/_ref, built by oracle/build_ref.py) behind the same module
With a third argument "decoder" the whole pixel decoder (msdeformattn.py:169-368: input projections + GroupNorm, position
encoding, the encoder, the FPN top-down pass and the mask-feature head; COMBO-R50 S4 settings) runs on ResNet-50-shaped synthetic
backbone features of a 224x224 frame (res2 256x56x56 ... res5 2048x7x7) instead of the bare encoder.
Under torchrun (WORLD_SIZE > 1) the encoder is wrapped in DistributedDataParallel: one process per GPU, every rank its own
frames (weak scaling), NCCL only for the gradient all-reduce.  Prints one JSON object (rank 0)."""
import json, os, statistics, sys, time
os.environ["NCCL_DEBUG_FILE"] = os.environ.get("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import combo_avs_b200 as pkg
from combo_avs_b200 import modules as pkg_modules
from combo_avs_b200.encoder import MSDeformAttnTransformerEncoderOnly
from combo_avs_b200.pixel_decoder import MSDeformAttnPixelDecoder

S4 = [(7, 7), (14, 14), (28, 28)]


class _RefFunction(torch.autograd.Function):
    """MSDeformAttnFunction over the reference extension (functions/ms_deform_attn_func.py:32-50)."""
    ext = None

    @staticmethod
    def forward(ctx, value, shapes, start, loc, attn, step):
        ctx.step = step
        out = _RefFunction.ext.ms_deform_attn_forward(value, shapes, start, loc, attn, step)
        ctx.save_for_backward(value, shapes, start, loc, attn)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        value, shapes, start, loc, attn = ctx.saved_tensors
        gv, gl, ga = _RefFunction.ext.ms_deform_attn_backward(value, shapes, start, loc, attn, grad_output.contiguous(), ctx.step)
        return gv, None, None, gl, ga, None


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    decoder = len(sys.argv) > 3 and sys.argv[3] == "decoder"
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(100 + rank)
    srcs = [torch.randn(frames, 256, h, w, generator=g).to(dev) for h, w in S4]
    poss = [torch.randn(frames, 256, h, w, generator=g).to(dev) * 0.5 for h, w in S4]
    gmem = torch.randn(frames, sum(h * w for h, w in S4), 256, generator=g).to(dev)
    r50 = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}
    feats = {k: torch.randn(frames, c, 224 // s, 224 // s, generator=g).to(dev) for k, (c, s) in r50.items()} if decoder else None
    res = {"workload": "pixel decoder (R50-shaped features, 224x224)" if decoder else "deformable encoder", "frames_per_gpu": frames,
           "n_gpus": world, "layers": 6, "dtype": "f32", "arms": {}}
    import build_ref
    ref_ext = build_ref.load_ref()
    arms = [("fused", True, None), ("unfused", False, None)] + ([("reference", False, ref_ext)] if ref_ext is not None else [])
    arms.append(("fused_tf32_matmul", True, None))  # information only: NOT the reference's numerics (it runs strict-fp32 GEMMs)
    if world == 1:  # the whole training step (forward, backward, AdamW) captured once in a CUDA graph and replayed
        arms.append(("fused_graph", True, None))
    original = pkg_modules.MSDeformAttnFunction
    for name, fuse, ext in arms:
        torch.manual_seed(7)
        if decoder:
            top = MSDeformAttnPixelDecoder(r50, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                           transformer_enc_layers=6, conv_dim=256, mask_dim=256, norm="GN",
                                           transformer_in_features=["res3", "res4", "res5"], common_stride=4).to(dev)
            enc = top.transformer
        else:
            top = enc = MSDeformAttnTransformerEncoderOnly(d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.0,
                                                           num_feature_levels=3, enc_n_points=4).to(dev)
        with torch.no_grad():  # off the zero init, so offsets / weights vary per query as in a trained model
            for layer in enc.encoder.layers:
                layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
                layer.self_attn.attention_weights.weight.normal_(0, 0.1)
                layer.self_attn.fuse_prologue = fuse
        torch.backends.cuda.matmul.allow_tf32 = name.endswith("tf32_matmul")
        _RefFunction.ext = ext
        pkg_modules.MSDeformAttnFunction = _RefFunction if ext is not None else original
        model = torch.nn.parallel.DistributedDataParallel(top, device_ids=[local]) if world > 1 else top
        graphed = name.endswith("_graph")
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=graphed)
        inputs = [x.clone().requires_grad_(True) for x in srcs]
        fin = {k: v.clone().requires_grad_(True) for k, v in feats.items()} if decoder else None

        def step():
            opt.zero_grad(set_to_none=True)
            if decoder:
                for x in fin.values():
                    x.grad = None
                mask_feat, low, multi = model(fin)
                (mask_feat.mean() + low.mean() + sum(m.mean() for m in multi)).backward()
            else:
                for x in inputs:
                    x.grad = None
                memory, _, _ = model(inputs, poss)
                memory.backward(gmem)
            opt.step()

        run = step
        if graphed:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            run = graph.replay
        for _ in range(5):
            run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        l0 = pkg.launch_count()
        for a, b in evs:
            a.record(); run(); b.record()
        torch.cuda.synchronize()
        ms = statistics.median(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        res["arms"][name] = {"ms_per_step": ms, "frames_per_s": world * frames / (ms * 1e-3), "msda_launches_per_step": (pkg.launch_count() - l0) / iters}
        if graphed:
            del graph, run
        del model, enc, top, opt
    pkg_modules.MSDeformAttnFunction = original
    torch.backends.cuda.matmul.allow_tf32 = False
    if "reference" in res["arms"]:
        res["speedup_vs_reference_extension"] = res["arms"]["reference"]["ms_per_step"] / res["arms"]["fused"]["ms_per_step"]
    if rank == 0:
        print(json.dumps(res, indent=1), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
