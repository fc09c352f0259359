"""Detectron2-free harness of the COMBO-AVS model AROUND the hot path -- measurement scaffolding, not product.

The product is combo-avs_b200/ (the MSDeformAttn core, its module and the pixel decoder).  Everything here exists so that BASELINE.json's
model-level configs (COMBO-R50 S4 forward on CPU; training step sharded by clip over 1-8 GPUs) can be RUN with the product on the hot path:
plain-PyTorch restatements of the reference's remaining modules (backbones with channel gate, VGGish, AVFuse, masked transformer decoder,
matcher, criterion, optimizer recipe) with random-init weights and synthetic AVSBench-shaped clips.  They follow the reference's structure
and tensor shapes (citations in each docstring) so that FLOPs, kernel mix and gradient all-reduce volume match; they are not parity-tested
bit for bit against the reference the way the product path is.
"""
