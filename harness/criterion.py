"""Matcher and criterion of COMBO (reference modeling/matcher.py:83-136, modeling/criterion.py:89-287) restated without detectron2,
with the two loss-side scaling blockers of SURVEY.md 8 f4 removed:
  * the reference computes each frame's matching cost in its own Python iteration and moves every cost matrix to the host separately
    (matcher.py:91-134: one .cpu() synchronisation per frame per decoder layer, 10 layers);  here ALL frames of ALL 10 prediction sets
    are costed in one batched pass on the GPU and cross to the host in ONE copy, then scipy solves the tiny assignments;
  * point sampling (detectron2's point_sample / get_uncertain_point_coords_with_randomness) is written with F.grid_sample directly.
Loss definitions, weights and the first-frame selection for S4 (criterion.py:241-254) are the reference's."""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment
from torch import nn


def point_sample(inp, coords):
    """detectron2 point_sample: inp [R, C, H, W], coords [R, P, 2] in [0, 1] (x, y) -> [R, C, P] (bilinear, align_corners False)."""
    return F.grid_sample(inp, 2.0 * coords.unsqueeze(2) - 1.0, align_corners=False).squeeze(3)


def uncertain_point_coords(logits, num_points, oversample_ratio, importance_sample_ratio):
    """get_uncertain_point_coords_with_randomness with uncertainty = -|logit| (criterion.py:70-86)."""
    R = logits.shape[0]
    n_sampled = int(num_points * oversample_ratio)
    coords = torch.rand(R, n_sampled, 2, device=logits.device)
    unc = -point_sample(logits, coords).abs()[:, 0]
    n_unc = int(importance_sample_ratio * num_points)
    idx = unc.topk(n_unc, dim=1)[1]
    picked = torch.gather(coords, 1, idx.unsqueeze(-1).expand(-1, -1, 2))
    if num_points > n_unc:
        picked = torch.cat((picked, torch.rand(R, num_points - n_unc, 2, device=logits.device)), 1)
    return picked


def dice_loss(inputs, targets, num_masks):
    inputs = inputs.sigmoid()
    num = 2 * (inputs * targets).sum(-1)
    den = inputs.sum(-1) + targets.sum(-1)
    return (1 - (num + 1) / (den + 1)).sum() / num_masks


def sigmoid_ce_loss(inputs, targets, num_masks):
    return F.binary_cross_entropy_with_logits(inputs, targets, reduction="none").mean(1).sum() / num_masks


class BatchedHungarianMatcher(nn.Module):
    """HungarianMatcher (matcher.py:66-136): cost = 2 * (-p[class]) + 5 * sigmoid-CE + 5 * dice over `num_points` random points shared
    by all masks of a frame.  forward() takes a LIST of prediction sets (last layer + auxiliary layers) and matches them all at once."""

    def __init__(self, cost_class=2.0, cost_mask=5.0, cost_dice=5.0, num_points=12544):
        super().__init__()
        self.cost_class, self.cost_mask, self.cost_dice, self.num_points = cost_class, cost_mask, cost_dice, num_points

    @torch.no_grad()
    def forward(self, prediction_sets, targets):
        n_sets, bs = len(prediction_sets), len(targets)
        logits = torch.stack([p["pred_logits"] for p in prediction_sets]).float()            # [S, B, Q, K+1]
        masks = torch.stack([p["pred_masks"] for p in prediction_sets]).float()              # [S, B, Q, h, w]
        Q = logits.shape[2]
        sizes = [len(t["labels"]) for t in targets]
        tmax = max(sizes) if sizes else 0
        if tmax == 0:
            empty = torch.zeros(0, dtype=torch.int64)
            return [[(empty, empty) for _ in range(bs)] for _ in range(n_sets)]
        coords = torch.rand(bs, self.num_points, 2, device=masks.device)                      # one point set per frame, shared by all layers
        out_pts = point_sample(masks.permute(1, 0, 2, 3, 4).flatten(1, 2), coords)            # [B, S*Q, P]
        out_pts = out_pts.view(bs, n_sets, Q, -1)
        tgt = torch.zeros(bs, tmax, *targets[0]["masks"].shape[-2:], device=masks.device)
        lab = torch.zeros(bs, tmax, dtype=torch.int64, device=masks.device)
        for b, t in enumerate(targets):
            tgt[b, :sizes[b]] = t["masks"].to(tgt)
            lab[b, :sizes[b]] = t["labels"]
        tgt_pts = point_sample(tgt, coords)                                                   # [B, T, P]
        prob = logits.softmax(-1).permute(1, 0, 2, 3)                                         # [B, S, Q, K+1]
        cost_class = -torch.gather(prob, 3, lab[:, None, None, :].expand(-1, n_sets, Q, -1))  # [B, S, Q, T]
        pos = F.binary_cross_entropy_with_logits(out_pts, torch.ones_like(out_pts), reduction="none")
        neg = F.binary_cross_entropy_with_logits(out_pts, torch.zeros_like(out_pts), reduction="none")
        P = out_pts.shape[-1]
        cost_mask = (torch.einsum("bsqp,btp->bsqt", pos, tgt_pts) + torch.einsum("bsqp,btp->bsqt", neg, 1 - tgt_pts)) / P
        sig = out_pts.sigmoid()
        num = 2 * torch.einsum("bsqp,btp->bsqt", sig, tgt_pts)
        den = sig.sum(-1)[..., None] + tgt_pts.sum(-1)[:, None, None, :]
        cost_dice = 1 - (num + 1) / (den + 1)
        C = (self.cost_mask * cost_mask + self.cost_class * cost_class + self.cost_dice * cost_dice).cpu()  # the ONE host copy
        result = []
        for s in range(n_sets):
            per_set = []
            for b in range(bs):
                i, j = linear_sum_assignment(C[b, s, :, :sizes[b]])
                per_set.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
            result.append(per_set)
        return result


class SetCriterion(nn.Module):
    """criterion.py:89-287: weighted cross-entropy on the matched classes (no-object weight 0.1), point-sampled sigmoid-CE + dice on the
    matched masks, for the last and the 9 auxiliary prediction sets, plus the frame-to-frame cosine loss on the per-layer mask logits."""

    def __init__(self, num_classes=2, eos_coef=0.1, num_points=12544, oversample_ratio=3.0, importance_sample_ratio=0.75,
                 weights=(2.0, 5.0, 5.0, 10.0), num_frames=5):
        super().__init__()
        self.num_classes, self.num_points = num_classes, num_points
        self.oversample_ratio, self.importance_sample_ratio, self.num_frames = oversample_ratio, importance_sample_ratio, num_frames
        self.w_ce, self.w_mask, self.w_dice, self.w_cos = weights
        self.matcher = BatchedHungarianMatcher(self.w_ce, self.w_mask, self.w_dice, num_points)
        empty_weight = torch.ones(num_classes + 1)
        empty_weight[-1] = eos_coef
        self.register_buffer("empty_weight", empty_weight)

    def _losses(self, pred, targets, indices, num_masks):
        logits = pred["pred_logits"].float()
        batch_idx = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        src_idx = torch.cat([src for src, _ in indices])
        target_classes = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        target_classes[batch_idx, src_idx] = torch.cat([t["labels"][j] for t, (_, j) in zip(targets, indices)])
        loss_ce = F.cross_entropy(logits.transpose(1, 2), target_classes, self.empty_weight)
        src_masks = pred["pred_masks"][batch_idx, src_idx][:, None]
        tgt_masks = torch.cat([t["masks"][j] for t, (_, j) in zip(targets, indices)]).to(src_masks)[:, None]
        with torch.no_grad():
            coords = uncertain_point_coords(src_masks, self.num_points, self.oversample_ratio, self.importance_sample_ratio)
            point_labels = point_sample(tgt_masks, coords).squeeze(1)
        point_logits = point_sample(src_masks, coords).squeeze(1)
        return loss_ce, sigmoid_ce_loss(point_logits, point_labels, num_masks), dice_loss(point_logits, point_labels, num_masks)

    def _cosine(self, middle):
        bs = middle.shape[0] // self.num_frames
        m = middle.reshape(bs, self.num_frames, -1)
        loss = 0
        for f in range(self.num_frames - 1):
            c = 1 - F.cosine_similarity(m[:, f], m[:, f + 1], dim=1, eps=1e-8)  # CosineEmbeddingLoss with target +1
            loss = loss + c * torch.exp(-c)
        return loss.sum() / bs / (self.num_frames - 1)

    def forward(self, outputs, targets):
        sets = [{"pred_logits": outputs["pred_logits"], "pred_masks": outputs["pred_masks"]}] + list(outputs["aux_outputs"])
        if len(sets[0]["pred_logits"]) != len(targets):  # S4: only the first frame of every clip is annotated (criterion.py:241-254)
            first = torch.arange(0, len(sets[0]["pred_logits"]), self.num_frames, device=sets[0]["pred_logits"].device)
            sets = [{k: v.index_select(0, first) for k, v in s.items()} for s in sets]
        indices = self.matcher(sets, targets)
        num_masks = torch.as_tensor([sum(len(t["labels"]) for t in targets)], dtype=torch.float, device=sets[0]["pred_logits"].device)
        world = 1
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(num_masks)  # criterion.py:263-265
            world = dist.get_world_size()
        num_masks = torch.clamp(num_masks / world, min=1)
        total = 0
        for s, idx in zip(sets, indices):
            ce, mk, dc = self._losses(s, targets, idx, num_masks)
            total = total + self.w_ce * ce + self.w_mask * mk + self.w_dice * dc
        for middle in outputs.get("middles_attn_mask", []):
            total = total + self.w_cos * self._cosine(middle)
        return total
