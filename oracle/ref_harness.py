"""TEST INFRASTRUCTURE ONLY - loader that runs the *unmodified* reference hot path.

The reference (``/root/reference``, read-only) is pure Python but cannot be
imported as-is in this image: ``model/cam_encoder.py:4`` needs
``efficientnet_pytorch``, ``model/convolutions.py:7`` needs ``timm`` and
``tool/geometry.py:6`` needs ``pyquaternion``; none is installed and there is no
network.  None of the three is used by the lift-splat path itself, so this
module registers inert stand-ins for them in ``sys.modules``, imports
``model.bev_model`` from the reference tree where it lies, swaps ``CamEncoder``
for a fake that returns preset tensors, and (on a CPU-only host) turns the
reference's hard-coded ``.cuda()`` calls (``model/bev_model.py:46,53``) into
no-ops.  Nothing of the reference is copied; nothing here is imported by the
product package.

Used by ``tests/golden/make_golden.py`` (to freeze reference outputs as
fixtures) and by the ``not gpu`` oracle-pinning tests when the reference tree
is present.  It does not exist on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

# /root/reference in the build container; baseline/_ref/ is where a driver-provided copy of the
# reference would sit on the GPU box (git-ignored; this repo never writes it)
REFERENCE_ROOTS = ("/root/reference",
                   os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"))


def reference_root():
    for r in REFERENCE_ROOTS:
        if os.path.isfile(os.path.join(r, "model", "bev_model.py")):
            return r
    return None


def available() -> bool:
    return reference_root() is not None


class _PresetEncoder(nn.Module):
    """Stands in for CamEncoder: returns whatever ``preset`` holds."""

    def __init__(self, *a, **k):
        super().__init__()
        self.preset = None

    def forward(self, images):
        feat, logits = self.preset
        return feat, logits


_loaded = None


def load_reference():
    """Import the reference ``model.bev_model`` module with inert stubs."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (expected /root/reference)")

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("pyquaternion", Quaternion=object)
    stub("efficientnet_pytorch", EfficientNet=object)
    stub("timm")
    stub("timm.models")
    stub("timm.models.layers", DropPath=nn.Identity, trunc_normal_=nn.init.trunc_normal_)
    if root not in sys.path:
        sys.path.insert(0, root)
    import model.bev_model as bm  # noqa: the reference's own module
    bm.CamEncoder = _PresetEncoder
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # reference hard-codes .cuda()
    _loaded = bm
    return bm


def reference_bev_model(cfg):
    """Instantiate the reference BevModel on ``cfg`` (a Configuration look-alike)."""
    bm = load_reference()
    return bm.BevModel(cfg)


def run_reference(cfg, feat, logits, intrinsics, extrinsics, double=False, backward_with=None):
    """Run reference forward (and optionally backward) on preset encoder outputs.

    double=True feeds ``image_feature.double()`` into ``proj_bev_feature`` (SURVEY.md
    8c): cumsum in fp64, one rounding on the store into the fp32 output.
    Returns a dict of tensors.
    """
    model = reference_bev_model(cfg)
    b, n = intrinsics.shape[:2]
    feat = feat.detach().clone().requires_grad_(backward_with is not None)
    logits = logits.detach().clone().requires_grad_(backward_with is not None)
    model.cam_encoder.preset = (feat, logits)
    images = torch.zeros(b, n, 3, 1, 1)
    geom = model.get_geometry(intrinsics, extrinsics)
    x, prob = model.encoder_forward(images)
    if double:
        # the per-sample scratch grid (model/bev_model.py:101) is allocated with the
        # default dtype; make it float64 so the fp64 sums reach the one rounding on
        # the store into the explicitly-float32 ``output`` (model/bev_model.py:76,105)
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            bev = model.proj_bev_feature(geom, x.double())
        finally:
            torch.set_default_dtype(prev)
    else:
        bev = model.proj_bev_feature(geom, x)
    out = {"geom": geom.detach(), "bev": bev.detach(), "prob": prob.detach()}
    if backward_with is not None:
        gb, gp = backward_with
        loss = (bev * gb.to(bev.dtype)).sum()
        if gp is not None:
            loss = loss + (prob * gp).sum()
        loss.backward()
        out["grad_feat"] = feat.grad.detach()
        out["grad_logits"] = logits.grad.detach()
    return out


def reference_stepper(cfg):
    """A callable running one forward+backward of the unmodified reference BevModel on preset
    encoder outputs (what bench.py times as kind "reference" when the tree is present):
    step(feat, logits, intrinsics, extrinsics, grad_bev, grad_prob) -> (bev, prob, gfeat, glogits)."""
    model = reference_bev_model(cfg)

    def step(feat, logits, intrinsics, extrinsics, grad_bev, grad_prob):
        model.to(feat.device)
        b, n = intrinsics.shape[:2]
        f = feat.detach().requires_grad_(True)
        z = logits.detach().requires_grad_(True)
        model.cam_encoder.preset = (f, z)
        images = torch.zeros(b, n, 3, 1, 1, device=feat.device)
        bev, prob = model(images, intrinsics, extrinsics)
        torch.autograd.backward([bev, prob], [grad_bev, grad_prob])
        return bev.detach(), prob.detach(), f.grad, z.grad

    return step


def reference_indices(cfg, intrinsics, extrinsics):
    """Voxel index (i64[B,Npts,3]), keep mask and sorted ranks, following the
    reference's own tensor ops line by line (model/bev_model.py:85-97) on the
    reference's own geometry."""
    model = reference_bev_model(cfg)
    geom = model.get_geometry(intrinsics, extrinsics)
    b = geom.shape[0]
    npts = geom[0].numel() // 3
    vox, keep, ranks = [], [], []
    for i in range(b):
        g = ((geom[i] - (model.bev_start_pos - model.bev_res / 2.0)) / model.bev_res)
        g = g.view(npts, 3).long()
        m = ((g[:, 0] >= 0) & (g[:, 0] < model.bev_dim[0])
             & (g[:, 1] >= 0) & (g[:, 1] < model.bev_dim[1])
             & (g[:, 2] >= 0) & (g[:, 2] < model.bev_dim[2]))
        gk = g[m]
        r = (gk[:, 0] * (model.bev_dim[1] * model.bev_dim[2]) + gk[:, 1] * model.bev_dim[2]) + gk[:, 2]
        vox.append(g)
        keep.append(m)
        ranks.append(r[r.argsort()])
    return geom, torch.stack(vox), torch.stack(keep), ranks
