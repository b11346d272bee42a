"""Stock-PyTorch stand-in for the reference's ParkingModel around the B200 lift-splat.

The reference stack cannot be imported on the GPU box (no reference tree there; here it needs
efficientnet_pytorch / timm / pytorch_lightning, all absent), so the sub-networks the
north_star keeps "in stock PyTorch" are rebuilt from torch / torchvision modules with the
reference's tensor shapes and (approximately) parameter counts:

  camera encoder  torchvision efficientnet_b4 trunk up to stride 16 (taps: 56 ch @ s8, 160 ch @ s16,
                  the reference's reduction_3 / reduction_4, model/cam_encoder.py:20,88-89) + two
                  atrous-pyramid heads with an up-sample-and-concat stage each -> feat[B*N,64,32,32],
                  depth_logits[B*N,48,32,32]
  lift-splat      e2e_parking_carla_b200.BevModel   (the product: sm_100a kernels)
  target channel  e2e_parking_carla_b200.add_target_bev
  BEV encoder     bilinear 200->256, ResNet-18 conv1..layer3 on 65 channels (model/bev_encoder.py:10-36;
                  its never-called layer4 is left out so that DDP needs no unused-parameter search)
  fusion          4-layer TransformerEncoder d=258 h=6 over 256 BEV tokens + ego-motion MLP
  control         4-layer TransformerDecoder, 204-token vocabulary, 14-step target
  segmentation    1x1-conv top-down path 16->32->64->128->200, 3 classes
  losses          token cross-entropy (PAD ignored), class-weighted segmentation cross-entropy,
                  e2e_parking_carla_b200.DepthLoss          (trainer/pl_trainer.py:55-83)

``lift_splat="torch"`` swaps the product for the reference's own torch op chain
(oracle/torch_port.py) - only bench.py's reference arm uses it.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import nn


def default_cfg(device="cuda"):
    """The fields of config/training.yaml the model reads (tool/config.py)."""
    return SimpleNamespace(
        device=device, token_nums=204, bev_encoder_in_channel=64, bev_encoder_out_channel=258,
        bev_x_bound=[-10.0, 10.0, 0.1], bev_y_bound=[-10.0, 10.0, 0.1], bev_z_bound=[-10.0, 10.0, 20.0],
        d_bound=[0.5, 12.5, 0.25], final_dim=[256, 256], bev_down_sample=8, use_depth_distribution=1,
        backbone="efficientnet-b4", seg_classes=3, seg_vehicle_weights=[1.0, 2.0, 2.0],
        tf_en_dim=258, tf_en_heads=6, tf_en_layers=4, tf_en_dropout=0.05, tf_en_bev_length=256,
        tf_en_motion_length=3, tf_de_dim=258, tf_de_heads=6, tf_de_layers=4, tf_de_dropout=0.05,
        tf_de_tgt_dim=15, learning_rate=1e-4, weight_decay=1e-4, batch_size=12)


# ----------------------------------------------------------------------------------------
# camera encoder stand-in
# ----------------------------------------------------------------------------------------
def _cbr(cin, cout, k=3, dilation=1):
    pad = dilation * (k // 2)
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=pad, dilation=dilation, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class AtrousPyramidHead(nn.Module):
    """1x1 + three dilated 3x3 + pooled branch -> 1x1 projection -> 3x3 -> 1x1 (DeepLab-v3 style)."""

    def __init__(self, cin, cout, hidden=64, rates=(12, 24, 36)):
        super().__init__()
        self.branches = nn.ModuleList([_cbr(cin, hidden, 1)] + [_cbr(cin, hidden, 3, r) for r in rates])
        self.pooled = nn.Sequential(nn.AdaptiveAvgPool2d(1), _cbr(cin, hidden, 1))
        self.project = nn.Sequential(_cbr(hidden * (len(rates) + 2), hidden, 1), nn.Dropout(0.5))
        self.tail = nn.Sequential(_cbr(hidden, hidden, 3), nn.Conv2d(hidden, cout, 1))

    def forward(self, x):
        outs = [b(x) for b in self.branches]
        outs.append(F.interpolate(self.pooled(x), size=x.shape[-2:], mode="bilinear", align_corners=False))
        return self.tail(self.project(torch.cat(outs, dim=1)))


class UpFuse(nn.Module):
    """x2 bilinear up-sample, concat with the skip tap, two 3x3 conv-BN-ReLU (ends in ReLU)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.body = nn.Sequential(_cbr(cin, cout, 3), _cbr(cout, cout, 3))

    def forward(self, deep, skip):
        deep = F.interpolate(deep, scale_factor=2, mode="bilinear", align_corners=False)
        return self.body(torch.cat([skip, deep], dim=1))


class StandInCamEncoder(nn.Module):
    def __init__(self, cfg, depth_bins):
        super().__init__()
        from torchvision.models import efficientnet_b4
        trunk = efficientnet_b4(weights=None).features
        self.to_s8 = trunk[:4]          # stem + stages 1-3 -> 56 ch @ stride 8
        self.to_s16 = trunk[4:6]        # stages 4-5      -> 160 ch @ stride 16
        c8, c16 = 56, 160
        self.feat_head = AtrousPyramidHead(c16, c16)
        self.feat_fuse = UpFuse(c16 + c8, cfg.bev_encoder_in_channel)
        self.depth_head = AtrousPyramidHead(c16, c16)
        self.depth_fuse = UpFuse(c16 + c8, depth_bins)

    def forward(self, x):
        s8 = self.to_s8(x)
        s16 = self.to_s16(s8)
        return self.feat_fuse(self.feat_head(s16), s8), self.depth_fuse(self.depth_head(s16), s8)


# ----------------------------------------------------------------------------------------
# the reference's lift-splat as torch ops (reference arm only)
# ----------------------------------------------------------------------------------------
class TorchOpsBevModel(nn.Module):
    def __init__(self, cfg, cam_encoder):
        super().__init__()
        from oracle import lift_splat_oracle as lo
        res, start, dim = lo.bev_grid_params(cfg.bev_x_bound, cfg.bev_y_bound, cfg.bev_z_bound)
        self.register_buffer("bev_res", torch.from_numpy(res))
        self.register_buffer("bev_start_pos", torch.from_numpy(start))
        self.bev_dim = [int(v) for v in dim]
        self.register_buffer("frustum", torch.from_numpy(lo.create_frustum(cfg.d_bound, cfg.final_dim,
                                                                          cfg.bev_down_sample)))
        self.cam_encoder = cam_encoder

    def forward(self, images, intrinsics, extrinsics):
        from oracle import torch_port as tp
        b, n, c, h, w = images.shape
        feat, logits = self.cam_encoder(images.view(b * n, c, h, w))
        bev, prob = tp.lift_splat_cpu(feat, logits, intrinsics, extrinsics, self.frustum, self.bev_start_pos,
                                      self.bev_res, self.bev_dim)
        return bev, prob


# ----------------------------------------------------------------------------------------
# downstream of the BEV
# ----------------------------------------------------------------------------------------
class BevEncoderStandIn(nn.Module):
    def __init__(self, in_channel):
        super().__init__()
        from torchvision.models.resnet import resnet18
        r = resnet18(weights=None, zero_init_residual=True)
        self.stem = nn.Sequential(nn.Conv2d(in_channel + 1, 64, 7, stride=2, padding=3, bias=False), r.bn1, r.relu,
                                  r.maxpool)
        self.stages = nn.Sequential(r.layer1, r.layer2, r.layer3)

    def forward(self, x):
        x = F.interpolate(x, size=(256, 256), mode="bilinear", align_corners=False)
        return torch.flatten(self.stages(self.stem(x)), 2)            # [B, 256, 16*16]


class FusionStandIn(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=cfg.tf_en_dim, nhead=cfg.tf_en_heads)
        self.encoder = nn.TransformerEncoder(layer, num_layers=cfg.tf_en_layers, enable_nested_tensor=False)
        n = cfg.tf_en_bev_length
        self.pos = nn.Parameter(torch.randn(1, n, cfg.tf_en_dim) * 0.02)
        self.drop = nn.Dropout(cfg.tf_en_dropout)
        self.motion = nn.Sequential(nn.Linear(cfg.tf_en_motion_length, n // 4), nn.ReLU(inplace=True),
                                    nn.Linear(n // 4, n // 2), nn.ReLU(inplace=True),
                                    nn.Linear(n // 2, n), nn.ReLU(inplace=True))

    def forward(self, bev_tokens, ego_motion):
        tokens = bev_tokens.transpose(1, 2)                                   # [B, 256 tokens, 256 ch]
        motion = self.motion(ego_motion).transpose(1, 2).expand(-1, -1, 2)    # [B, 256, 2]
        x = self.drop(torch.cat([tokens, motion], dim=2) + self.pos)
        return self.encoder(x.transpose(0, 1)).transpose(0, 1)               # [B, 256, 258]


class ControlStandIn(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.pad = cfg.token_nums - 1
        self.embed = nn.Embedding(cfg.token_nums, cfg.tf_de_dim)
        self.pos = nn.Parameter(torch.randn(1, cfg.tf_de_tgt_dim - 1, cfg.tf_de_dim) * 0.02)
        self.drop = nn.Dropout(cfg.tf_de_dropout)
        layer = nn.TransformerDecoderLayer(d_model=cfg.tf_de_dim, nhead=cfg.tf_de_heads)
        self.decoder = nn.TransformerDecoder(layer, num_layers=cfg.tf_de_layers)
        self.out = nn.Linear(cfg.tf_de_dim, cfg.token_nums)

    def _decode(self, memory, tgt, emb):
        L = tgt.shape[1]
        causal = torch.full((L, L), float("-inf"), device=tgt.device).triu(1)
        # tgt_is_causal=False: use the mask as given; the default (None) makes torch compare it with a
        # causal mask on the host (a device->host sync per call, and not capturable in a CUDA graph)
        y = self.decoder(tgt=emb.transpose(0, 1), memory=memory.transpose(0, 1), tgt_mask=causal,
                         tgt_key_padding_mask=(tgt == self.pad), tgt_is_causal=False)
        return self.out(y.transpose(0, 1))

    def forward(self, memory, tgt):
        tgt = tgt[:, :-1]
        return self._decode(memory, tgt, self.drop(self.embed(tgt) + self.pos))

    def predict(self, memory, tgt):
        n = tgt.shape[1]
        pad = torch.full((tgt.shape[0], self.cfg.tf_de_tgt_dim - n - 1), self.pad, dtype=torch.long, device=tgt.device)
        full = torch.cat([tgt, pad], dim=1)
        logits = self._decode(memory, full, self.embed(full) + self.pos)[:, n - 1]
        return logits.softmax(dim=-1).argmax(dim=-1).view(-1, 1)


class SegHeadStandIn(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        cin, c = cfg.bev_encoder_out_channel, cfg.bev_encoder_in_channel
        self.lateral = nn.ModuleList([nn.Conv2d(cin, c, 1)] + [nn.Conv2d(c, c, 1) for _ in range(3)])
        self.head = nn.Sequential(nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c), nn.ReLU(inplace=True),
                                  nn.Conv2d(c, cfg.seg_classes, 1))

    def forward(self, fused):
        b, s, c = fused.shape
        side = int(math.sqrt(s))
        x = fused.transpose(1, 2).reshape(b, c, side, side)
        x = F.relu(self.lateral[0](x))
        for conv in self.lateral[1:]:
            x = F.relu(conv(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)))
        return self.head(F.interpolate(x, size=(200, 200), mode="bilinear", align_corners=False))


# ----------------------------------------------------------------------------------------
# the whole model + one training step
# ----------------------------------------------------------------------------------------
class ParkingStack(nn.Module):
    """forward(batch) -> (pred_control, pred_segmentation, pred_depth)   model/parking_model.py:67-70
    predict(batch)    -> (tokens, pred_segmentation, pred_depth, bev_target)          :72-78"""

    def __init__(self, cfg, lift_splat="b200", channels_last=True):
        super().__init__()
        self.cfg = cfg
        depth_bins = int(torch.arange(*cfg.d_bound, dtype=torch.float).numel())
        enc = StandInCamEncoder(cfg, depth_bins)
        self.native = lift_splat == "b200"
        if self.native:
            from e2e_parking_carla_b200 import BevModel
            if channels_last:
                enc = enc.to(memory_format=torch.channels_last)
            self.bev_model = BevModel(cfg, cam_encoder=enc, spare_channels=1 if channels_last else 0,
                                      bev_memory_format=torch.channels_last if channels_last else torch.contiguous_format)
        else:
            self.bev_model = TorchOpsBevModel(cfg, enc)
        self.channels_last = channels_last and self.native
        self.bev_encoder = BevEncoderStandIn(cfg.bev_encoder_in_channel)
        if self.channels_last:
            self.bev_encoder = self.bev_encoder.to(memory_format=torch.channels_last)
        self.fusion = FusionStandIn(cfg)
        self.control = ControlStandIn(cfg)
        self.seg_head = SegHeadStandIn(cfg)

    def add_target(self, bev, target_point):
        if self.native:
            from e2e_parking_carla_b200 import add_target_bev
            return add_target_bev(bev, target_point, self.cfg, noise=True)
        from oracle import torch_port as tp
        return tp.add_target_bev_ref(bev, target_point, self.cfg.bev_x_bound[2], self.cfg.bev_y_bound[2])

    def encoder(self, data):
        images = data["image"]
        bev, depth = self.bev_model(images, data["intrinsics"], data["extrinsics"])
        bev, target_map = self.add_target(bev, data["target_point"])
        fused = self.fusion(self.bev_encoder(bev), data["ego_motion"])
        return fused, self.seg_head(fused), depth, target_map

    def forward(self, data):
        fused, seg, depth, _ = self.encoder(data)
        return self.control(fused, data["gt_control"]), seg, depth

    @torch.no_grad()
    def predict(self, data):
        fused, seg, depth, target_map = self.encoder(data)
        tokens = data["gt_control"]
        for _ in range(3):
            tokens = torch.cat([tokens, self.control.predict(fused, tokens)], dim=1)
        return tokens, seg, depth, target_map


class Losses(nn.Module):
    """control + segmentation + depth, summed (trainer/pl_trainer.py:55-83)."""

    def __init__(self, cfg, native=True):
        super().__init__()
        self.pad = cfg.token_nums - 1
        self.register_buffer("seg_w", torch.tensor(cfg.seg_vehicle_weights, dtype=torch.float))
        self.native = native
        if native:
            from e2e_parking_carla_b200 import DepthLoss
            self.depth = DepthLoss(cfg)
        self.cfg = cfg

    def depth_torch(self, prob, gt):
        """loss/depth_loss.py:18-48 with torch ops (reference arm)."""
        cfg, ds = self.cfg, self.cfg.bev_down_sample
        D = prob.shape[1]
        bn = prob.shape[0]
        g = gt.reshape(bn, gt.shape[-2] // ds, ds, gt.shape[-1] // ds, ds).permute(0, 1, 3, 2, 4).reshape(-1, ds * ds)
        g = torch.where(g == 0.0, torch.full_like(g, 1e5), g).min(dim=-1).values
        g = (g - (cfg.d_bound[0] - cfg.d_bound[2])) / cfg.d_bound[2]
        g = torch.where((g < D + 1) & (g >= 0.0), g, torch.zeros_like(g))
        onehot = F.one_hot(g.long(), num_classes=D + 1)[:, 1:].float()
        p = prob.permute(0, 2, 3, 1).reshape(-1, D)
        fg = onehot.max(dim=1).values > 0
        return F.binary_cross_entropy(p[fg], onehot[fg], reduction="none").sum() / max(1.0, float(fg.sum()))

    def forward(self, pred, data):
        ctrl, seg, depth = pred
        l_ctrl = F.cross_entropy(ctrl.reshape(-1, ctrl.shape[-1]), data["gt_control"][:, 1:].reshape(-1),
                                 ignore_index=self.pad)
        l_seg = F.cross_entropy(seg, data["segmentation"].view(seg.shape[0], *seg.shape[-2:]), reduction="none",
                                ignore_index=255, weight=self.seg_w).mean()
        l_depth = self.depth(depth, data["depth"]) if self.native else self.depth_torch(depth, data["depth"])
        return l_ctrl + l_seg + l_depth


def synthetic_batch(cfg, batch, device, seed=0, cams=4):
    """Shapes of one CarlaDataset batch (dataset/carla_dataset.py:379-423)."""
    from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_depth_labels, make_rig
    g = torch.Generator().manual_seed(seed)
    h, w = cfg.final_dim
    intr, extr = make_rig(batch, cams, jitter=True, seed=seed + 1)
    tokens = torch.randint(0, 201, (batch, 12), generator=g)
    ctrl = torch.cat([torch.full((batch, 1), 201), tokens, torch.full((batch, 1), 202), torch.full((batch, 1), 203)], 1)
    shape = LiftSplatShape(batch=batch, cams=cams, final_dim=list(cfg.final_dim), bev_down_sample=cfg.bev_down_sample)
    data = {
        "image": torch.randn(batch, cams, 3, h, w, generator=g),
        "depth": make_depth_labels(shape, seed=seed),
        "intrinsics": intr, "extrinsics": extr,
        "segmentation": torch.randint(0, 3, (batch, 1, 200, 200), generator=g),
        "target_point": torch.cat([(torch.rand(batch, 2, generator=g) * 16 - 8), torch.zeros(batch, 1)], 1),
        "ego_motion": torch.randn(batch, 1, 3, generator=g),
        "gt_control": ctrl,
    }
    return {k: v.to(device) for k, v in data.items()}


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
