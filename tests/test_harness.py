"""The detectron2-free COMBO-R50 harness (harness/): CPU tests of the criterion / batched matcher and of one
training step with the reference's CPU core on the hot path (oracle/core_pytorch.py); a GPU test of the same step on the product kernels."""
import itertools
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _fake_outputs(frames, queries, h, w, layers=3, seed=0, requires_grad=True):
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g).requires_grad_(requires_grad)
    sets = [{"pred_logits": mk(frames, queries, 3), "pred_masks": mk(frames, queries, h, w)} for _ in range(layers)]
    return {"pred_logits": sets[-1]["pred_logits"], "pred_masks": sets[-1]["pred_masks"], "aux_outputs": sets[:-1],
            "middles_attn_mask": [s["pred_masks"].reshape(frames, queries, -1) for s in sets[:-1]]}


def test_batched_matcher_is_optimal_and_criterion_backpropagates():
    from harness.combo_model import synthetic_clips
    from harness.criterion import BatchedHungarianMatcher, SetCriterion
    clips, frames, Q = 2, 10, 6
    _, _, _, targets = synthetic_clips(clips, size=32, seed=1)
    out = _fake_outputs(frames, Q, 8, 8)
    crit = SetCriterion(num_points=64)
    loss = crit(out, targets)  # S4: 10 frames of predictions, 2 annotated first frames
    assert torch.isfinite(loss)
    loss.backward()
    assert out["pred_masks"].grad is not None and out["pred_masks"].grad[0].abs().sum() > 0
    assert out["pred_masks"].grad[1].abs().sum() == 0  # frame 1 of clip 0 is not supervised by the mask losses ...
    assert out["aux_outputs"][0]["pred_masks"].grad[1].abs().sum() > 0  # ... but the cosine loss reaches the intermediate layers

    # optimality of the batched assignment against brute force, with the matcher's own cost (fixed points via the seed)
    matcher = BatchedHungarianMatcher(num_points=32)
    sets = [{"pred_logits": out["pred_logits"][::5].detach(), "pred_masks": out["pred_masks"][::5].detach()}]
    torch.manual_seed(5)
    idx = matcher(sets, targets)[0]
    torch.manual_seed(5)
    coords = torch.rand(2, 32, 2)
    from harness.criterion import point_sample
    for b, (i, j) in enumerate(idx):
        assert sorted(j.tolist()) == [0, 1] and len(set(i.tolist())) == 2
        o = point_sample(sets[0]["pred_masks"][b][None], coords[b][None])[0]
        t = point_sample(targets[b]["masks"].float()[None], coords[b][None])[0]
        prob = sets[0]["pred_logits"][b].softmax(-1)
        cost = np.zeros((Q, 2))
        for q, k in itertools.product(range(Q), range(2)):
            ce = torch.nn.functional.binary_cross_entropy_with_logits(o[q], t[k])
            s = o[q].sigmoid()
            dice = 1 - (2 * (s * t[k]).sum() + 1) / (s.sum() + t[k].sum() + 1)
            cost[q, k] = 5 * ce + 5 * dice - 2 * prob[q, targets[b]["labels"][k]]
        best = min(cost[a, 0] + cost[c, 1] for a, c in itertools.permutations(range(Q), 2))
        got = sum(cost[a, c] for a, c in zip(i.tolist(), j.tolist()))
        assert abs(got - best) < 1e-4


def _cpu_reference_core():
    """Route the module's op to the reference's CPU implementation (restated in oracle/core_pytorch.py) -- test-only."""
    import oracle.core_pytorch as cp
    from combo_avs_b200 import modules

    class F:
        @staticmethod
        def apply(value, shapes, start, loc, attn, step):
            return cp.msda_core_pytorch(value, shapes, loc, attn)
    old = modules.MSDeformAttnFunction
    modules.MSDeformAttnFunction = F
    return modules, old


def test_combo_r50_training_step_runs_on_cpu_with_the_reference_core():
    from harness.combo_model import ComboR50, build_optimizer, clip_full_model_gradients, synthetic_clips
    from harness.criterion import SetCriterion
    modules, old = _cpu_reference_core()
    try:
        torch.manual_seed(0)
        model = ComboR50().train()
        crit = SetCriterion(num_points=256)
        opt = build_optimizer(model)
        images, pre_masks, audio, targets = synthetic_clips(1, size=64, seed=2)
        out = model(images, pre_masks, audio)
        assert out["pred_masks"].shape == (5, 100, 16, 16) and out["pred_logits"].shape == (5, 100, 3) and len(out["aux_outputs"]) == 9
        loss = crit(out, targets)
        loss.backward()
        norm = clip_full_model_gradients(model, 0.01)
        opt.step()
        assert torch.isfinite(loss) and torch.isfinite(norm)
        assert all(p.grad is None for p in model.audio_backbone.parameters())          # VGGish is frozen
        assert model.pixel_decoder.transformer.encoder.layers[0].self_attn.sampling_offsets.weight.grad is not None
        assert len(opt.param_groups) <= 4                                              # grouped by hyper-parameters, not per parameter
    finally:
        modules.MSDeformAttnFunction = old


@pytest.mark.gpu
def test_combo_r50_training_step_runs_on_the_product_kernels():
    import combo_avs_b200 as pkg
    from harness.combo_model import ComboR50, build_optimizer, clip_full_model_gradients, synthetic_clips
    from harness.criterion import SetCriterion
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = ComboR50().to(dev).train()
    crit = SetCriterion().to(dev)
    opt = build_optimizer(model)
    images, pre_masks, audio, targets = synthetic_clips(2, device=dev, seed=3)
    before = pkg.launch_count()
    losses = []
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(images, pre_masks, audio), targets)
        loss.backward()
        clip_full_model_gradients(model, 0.01)
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses))
    assert pkg.launch_count() - before >= 2 * 6 * 2        # six deformable-attention layers, forward + backward, two steps
    assert pkg.last_variant("forward").startswith("slab32") and pkg.last_variant("backward").startswith("sorted32")


def test_pvtv2_b5_backbone_pyramid_and_avss_criterion_on_cpu():
    """BASELINE configs 4 / 5 (harness only): the PVTv2-B5 pyramid has the channels / strides the pixel decoder is built for
    (PRE_SAM_DIM [64, 128, 320, 512]), and the criterion takes the AVSS setting -- 71 classes, every frame of a 10-frame clip supervised."""
    from harness.combo_model import synthetic_clips
    from harness.criterion import SetCriterion
    from harness.pvtv2 import PVTV2_B5_FEATURES, PVTv2B5Features
    net = PVTv2B5Features(depths=(1, 1, 2, 1)).eval()  # (the full depths 3/6/40/3 only change the running time)
    with torch.no_grad():
        feats = net(torch.randn(2, 3, 64, 96))
    for name, (channels, stride) in PVTV2_B5_FEATURES.items():
        assert feats[name].shape == (2, channels, 64 // stride, 96 // stride), name
    assert sum(p.numel() for p in PVTv2B5Features().parameters()) == 81443008  # PVTv2-B5 without its classifier head
    clips, frames, Q, classes = 1, 10, 5, 71
    _, _, _, targets = synthetic_clips(clips, num_frames=frames, size=32, seed=2, supervised_frames=frames, num_classes=classes)
    assert len(targets) == frames and all(1 <= int(t["labels"][1]) < classes for t in targets)
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: torch.randn(*s, generator=g).requires_grad_(True)
    sets = [{"pred_logits": mk(frames, Q, classes + 1), "pred_masks": mk(frames, Q, 8, 8)} for _ in range(2)]
    out = {"pred_logits": sets[-1]["pred_logits"], "pred_masks": sets[-1]["pred_masks"], "aux_outputs": sets[:-1],
           "middles_attn_mask": [s["pred_masks"].reshape(frames, Q, -1) for s in sets[:-1]]}
    loss = SetCriterion(num_classes=classes, num_points=64, num_frames=frames)(out, targets)
    assert torch.isfinite(loss)
    loss.backward()
    assert all(out["pred_masks"].grad[f].abs().sum() > 0 for f in range(frames))  # every frame is supervised
