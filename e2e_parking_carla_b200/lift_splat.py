"""Host side of the lift-splat hot path: thin wrappers that hand torch device memory
and the current CUDA stream to the C ABI (``include/ls_b200.h``), plus the
``torch.autograd.Function`` that ``BevModel.calc_bev_feature`` routes through.

PyTorch is plumbing here (allocator, streams, autograd graph); every per-point
operation runs in ``libls_b200.so``.  CUDA tensors only - there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LS_BF16, LS_F32, LS_FEAT_NCHW, LS_FEAT_NHWC, LsBevStrides, LsShape, check


# --------------------------------------------------------------------------------------
# problem description
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class GridSpec:
    """bev_start_pos / bev_res / bev_dim of a BevModel as host scalars (read once at
    construction, so no call ever syncs on the on-device nn.Parameters the way
    model/bev_model.py:76,101 does)."""
    start: Tuple[float, float, float]
    res: Tuple[float, float, float]
    dim: Tuple[int, int, int]


def pick_tile_x(channels: int, bev_channels_last: bool) -> int:
    """Internal BEV tiling for a call (LsShape.tile_x).  1 x 128 strips are the default everywhere:
    the NCHW write-out and the bulk-store variant need them, and for the channels-last direct splat
    (which accepts any tile shape) square tiles measured SLOWER on a B200 although a ray stays 3.6
    depth bins inside an 8 x 16 tile against 1.15 inside a strip - that kernel is bound by L1 data-pipe
    wavefronts and issue slots, not by L2 reads (profiles/r02_summary.md).  LS_TILE_X=2..64 selects
    square-ish tiles for the 64-channel channels-last splat (developer knob, covered by the tests)."""
    import os
    if not (bev_channels_last and channels == 64) or os.environ.get("LS_SPLAT_OUT") == "bulk":
        return 1
    return int(os.environ.get("LS_TILE_X", "1"))


def make_shape(B: int, N: int, D: int, fh: int, fw: int, Cc: int, grid: GridSpec, geom_policy: int = 0,
               tile_x: int = 1, bev_dtype: int = LS_F32) -> LsShape:
    s = LsShape()
    s.geom_policy = geom_policy
    s.tile_x = tile_x
    s.bev_dtype = bev_dtype
    s.B, s.N, s.D, s.fh, s.fw, s.C = B, N, D, fh, fw, Cc
    s.X, s.Y, s.Z = grid.dim
    for i in range(3):
        s.start[i] = grid.start[i]
        s.res[i] = grid.res[i]
    return s


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return LS_F32
    if t.dtype == torch.bfloat16:
        return LS_BF16
    raise TypeError("lift-splat supports float32 and bfloat16, got %s" % t.dtype)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lift-splat kernels need CUDA tensors (no CPU fallback); got %s" % t.device)


def _bev_strides(t: torch.Tensor) -> LsBevStrides:
    """Element strides of a [B,C,X,Y] tensor; the library accepts NCHW-like (unit Y stride) and
    channels-last-like (unit C stride) tensors, including channel-slice views of either."""
    if t.dim() != 4:
        raise ValueError("BEV tensor must be [B,C,X,Y]")
    return LsBevStrides(t.stride(0), t.stride(1), t.stride(2), t.stride(3))


def _grad_layout_ok(g: torch.Tensor, depth_bins: int = 16) -> bool:
    """Can ls_backward read this gradient in place?  (mirrors ls_classify_grad_in)"""
    if g.dtype == torch.bfloat16:
        return (g.stride(1) == 1 and g.shape[1] == 64 and depth_bins % 16 == 0 and g.stride(2) == g.shape[3] * g.stride(3)
                and g.stride(3) >= 64 and g.stride(3) % 4 == 0 and g.stride(0) % 4 == 0 and g.data_ptr() % 8 == 0)
    if g.stride(3) == 1 and (g.stride(1) != 1 or g.shape[1] == 1):
        return True
    return (g.stride(1) == 1 and g.shape[1] % 4 == 0 and g.stride(2) == g.shape[3] * g.stride(3)
            and g.stride(3) >= g.shape[1] and g.shape[2] * g.shape[3] * g.stride(3) * 4 < 2 ** 32)


def _feat_layout(feat: torch.Tensor):
    """(tensor to hand to the library, LsFeatLayout): a channels_last feature map is consumed in
    place (no staging copy), everything else goes in as contiguous NCHW."""
    if feat.is_contiguous():
        return feat, LS_FEAT_NCHW
    if feat.shape[1] % 4 == 0 and feat.is_contiguous(memory_format=torch.channels_last):
        return feat, LS_FEAT_NHWC
    return feat.contiguous(), LS_FEAT_NCHW


# one transient scratch blob per (device, stream), grown on demand and reused by every call
_scratch = {}


def scratch_for(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def grid_cells(shape: LsShape) -> Tuple[int, int, int]:
    """(tiles, padded cell count, seg_start row stride) of the internal BEV tiling."""
    tiles, cells, stride = C.c_int32(), C.c_int32(), C.c_int32()
    check(_lib.load().ls_grid_cells(C.byref(shape), C.byref(tiles), C.byref(cells), C.byref(stride)),
          "ls_grid_cells")
    return tiles.value, cells.value, stride.value


def padded_channels(c: int) -> int:
    return int(_lib.load().ls_padded_channels(c))


# --------------------------------------------------------------------------------------
# one wrapper per ABI entry point (used by BevModel, the tests and bench.py)
# --------------------------------------------------------------------------------------
def camera_transform(intrinsics: torch.Tensor, extrinsics: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(M f32[B,N,3,3], t f32[B,N,3]) - model/bev_model.py:46-47,53."""
    _need_cuda(intrinsics, extrinsics)
    b, n = extrinsics.shape[:2]
    intr = intrinsics.detach().to(torch.float32).contiguous()
    extr = extrinsics.detach().to(torch.float32).contiguous()
    M = torch.empty(b, n, 3, 3, dtype=torch.float32, device=extr.device)
    t = torch.empty(b, n, 3, dtype=torch.float32, device=extr.device)
    check(_lib.load().ls_camera_transform(_ptr(intr), _ptr(extr), b * n, _ptr(M), _ptr(t), _stream(extr)),
          "ls_camera_transform")
    return M, t


def geometry(M, t, frustum, shape: LsShape) -> torch.Tensor:
    """geom f32[B,N,D,fh,fw,3] - model/bev_model.py:49-55 (compat / tests only)."""
    _need_cuda(M, t, frustum)
    geom = torch.empty(shape.B, shape.N, shape.D, shape.fh, shape.fw, 3, dtype=torch.float32, device=M.device)
    check(_lib.load().ls_geometry(_ptr(M.contiguous()), _ptr(t.contiguous()), _ptr(frustum.contiguous()),
                                  C.byref(shape), _ptr(geom), _stream(M)), "ls_geometry")
    return geom


def index(M, t, frustum, shape: LsShape, for_sort: bool = False):
    """rank i32[B,Npts] (-1 dropped); with ``for_sort`` also (cell, within, counts)."""
    _need_cuda(M, t, frustum)
    npts = shape.N * shape.D * shape.fh * shape.fw
    dev = M.device
    rank = torch.empty(shape.B, npts, dtype=torch.int32, device=dev)
    cell = within = counts = None
    if for_sort:
        cell = torch.empty_like(rank)
        within = torch.empty_like(rank)
        counts = torch.zeros(shape.B, grid_cells(shape)[1], dtype=torch.int32, device=dev)
    check(_lib.load().ls_index(_ptr(M.contiguous()), _ptr(t.contiguous()), _ptr(frustum.contiguous()),
                               C.byref(shape), _ptr(rank), _ptr(cell), _ptr(within), _ptr(counts), _stream(M)),
          "ls_index")
    return (rank, cell, within, counts) if for_sort else rank


def export_indices(M, t, frustum, shape: LsShape):
    """(vox i64[B,Npts,3], keep bool[B,Npts], rank i64[B,Npts]) in the reference's
    convention (model/bev_model.py:85-95) for the bit-exact tests."""
    _need_cuda(M, t, frustum)
    npts = shape.N * shape.D * shape.fh * shape.fw
    dev = M.device
    vox = torch.empty(shape.B, npts, 3, dtype=torch.int64, device=dev)
    keep = torch.empty(shape.B, npts, dtype=torch.uint8, device=dev)
    rank = torch.empty(shape.B, npts, dtype=torch.int64, device=dev)
    check(_lib.load().ls_export_indices(_ptr(M.contiguous()), _ptr(t.contiguous()), _ptr(frustum.contiguous()),
                                        C.byref(shape), _ptr(vox), _ptr(keep), _ptr(rank), _stream(M)),
          "ls_export_indices")
    return vox, keep.bool(), rank


def softmax(logits: torch.Tensor, shape: LsShape) -> torch.Tensor:
    _need_cuda(logits)
    logits = logits.contiguous()
    prob = torch.empty_like(logits)
    check(_lib.load().ls_softmax(_ptr(logits), _dtype_code(logits), C.byref(shape), _ptr(prob), _stream(logits)),
          "ls_softmax")
    return prob


def sort(cell, within, counts, prob, shape: LsShape, with_pixel_index: bool = False, parallel_scan: bool = True):
    """Counting sort by cell: (seg_start i32[B,seg_stride], tile_order i32[B,tiles], recs
    i32[B,Npts,2] = {key, prob bits}, pix_recs i32[B*N*HW, D, 2] or None)."""
    _need_cuda(cell, within, counts, prob)
    tiles, _, stride = grid_cells(shape)
    dev = cell.device
    npts = cell.shape[1]
    seg = torch.empty(shape.B, stride, dtype=torch.int32, device=dev)
    order = torch.empty(shape.B, tiles, dtype=torch.int32, device=dev)
    scratch = torch.empty(shape.B, tiles, dtype=torch.int32, device=dev) if parallel_scan else None
    recs = torch.zeros(shape.B, npts, 2, dtype=torch.int32, device=dev)
    pix = None
    if with_pixel_index:
        pix = torch.empty(shape.B * shape.N * shape.fh * shape.fw, shape.D, 2, dtype=torch.int32, device=dev)
    prob = prob.contiguous()
    check(_lib.load().ls_sort(_ptr(cell), _ptr(within), _ptr(counts), _ptr(prob), _dtype_code(prob), C.byref(shape),
                              _ptr(seg), _ptr(order), _ptr(scratch), _ptr(recs), _ptr(pix), _stream(cell)), "ls_sort")
    return seg, order, recs, pix


def export_sorted_ranks(seg_start: torch.Tensor, shape: LsShape, b: int) -> torch.Tensor:
    """``ranks[ranks.argsort()]`` of sample b (model/bev_model.py:97) rebuilt from the CSR."""
    cells = grid_cells(shape)[1]
    out = torch.empty(cells, 2, dtype=torch.int64, device=seg_start.device)
    check(_lib.load().ls_export_cell_counts(_ptr(seg_start), C.byref(shape), b, _ptr(out), None,
                                            _stream(seg_start)), "ls_export_cell_counts")
    out = out[out[:, 1] > 0]
    out = out[out[:, 0].argsort()]
    return torch.repeat_interleave(out[:, 0], out[:, 1])


def kept_counts(seg_start: torch.Tensor, shape: LsShape) -> torch.Tensor:
    kept = torch.empty(shape.B, dtype=torch.int32, device=seg_start.device)
    check(_lib.load().ls_export_cell_counts(_ptr(seg_start), C.byref(shape), 0, None, _ptr(kept),
                                            _stream(seg_start)), "ls_export_cell_counts")
    return kept


_UNSUPPORTED = ("unsupported lift-splat shape (need Z == 1, C <= 256, D <= 191, "
                "N*fh*fw*2^ceil(log2 D) <= 2^24; channels_last features need C % 4 == 0)")


def scratch_bytes(shape: LsShape, dtype_code: int, with_backward: bool) -> int:
    n = _lib.load().ls_scratch_bytes(C.byref(shape), dtype_code, int(with_backward))
    if n == 0:
        raise ValueError(_UNSUPPORTED)
    return n


def saved_bytes(shape: LsShape, dtype_code: int, feat_layout: int = LS_FEAT_NCHW) -> int:
    n = _lib.load().ls_saved_bytes(C.byref(shape), dtype_code, feat_layout)
    if n == 0:
        raise ValueError(_UNSUPPORTED)
    return n


class RigCache:
    """Opt-in cache of the index structures of a STATIC camera rig (ls_forward_cached): the CSR
    offsets, tile order, canonical records, slot -> point permutation and pixel-major index of one
    (shape, rig).  The reference's rig is a constant of the dataset (dataset/carla_dataset.py:392-393),
    so after the first step the index kernel, histogram, scan, placement and re-ordering are replaced
    by a streaming refresh of the record weights.  The owner vouches that intrinsics / extrinsics do
    not change between ``invalidate()`` calls; one forward/backward pair in flight per cache."""

    def __init__(self):
        self.blob, self.key, self.valid = None, None, False

    def invalidate(self):
        self.valid = False

    def prepare(self, shape: LsShape, code: int, device: torch.device) -> torch.Tensor:
        key = (shape.B, shape.N, shape.D, shape.fh, shape.fw, shape.C, shape.X, shape.Y, tuple(shape.start),
               tuple(shape.res), shape.geom_policy, shape.tile_x, shape.bev_dtype, code, device.index)
        if key != self.key or self.blob is None:
            n = _lib.load().ls_cache_bytes(C.byref(shape))
            if n == 0:
                raise ValueError(_UNSUPPORTED)
            self.blob, self.key, self.valid = torch.empty(n, dtype=torch.uint8, device=device), key, False
        return self.blob


# --------------------------------------------------------------------------------------
# autograd
# --------------------------------------------------------------------------------------
class LiftSplatFunction(torch.autograd.Function):
    """(feat, depth_logits) -> (bev, depth_prob); replaces softmax + outer product +
    voxelise + sort + VoxelsSumming + scatter (model/bev_model.py:64-105,
    tool/geometry.py:285-317) with one forward and one backward pipeline call.
    Gradients flow to ``feat`` and ``depth_logits`` only, as in the reference
    (geometry is cut by ``.long()``, bev_model.py:86).

    ``bev_format``: memory format of the returned ``bev`` ([B,C,X,Y] fp32 either way).
    ``torch.channels_last`` is the splat's native layout (a cell's channels are one row): the
    tile leaves as one bulk store, and a channels_last gradient is gathered in place."""

    # autocast: the function consumes whatever dtype the (autocast) encoder produced - bf16 feature maps
    # and logits go straight to the bf16 kernels, the BEV tensor is float32 either way as in the
    # reference (model/bev_model.py:76) - and runs with autocast switched off inside
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, feat, logits, M, t, frustum, shape: LsShape, bev_format=torch.contiguous_format,
                spare_channels: int = 0, rig_cache: Optional[RigCache] = None):
        _need_cuda(feat, logits, M, t, frustum)
        if feat.dtype != logits.dtype:
            raise TypeError("feat and depth logits must share a dtype")
        code = _dtype_code(feat)
        dev = feat.device
        feat_c, layout = _feat_layout(feat)
        logits_c = logits.contiguous()
        need_bwd = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        scratch = scratch_for(scratch_bytes(shape, code, need_bwd), dev)
        saved = torch.empty(saved_bytes(shape, code, layout), dtype=torch.uint8, device=dev) if need_bwd else None
        if spare_channels and bev_format != torch.channels_last:
            raise ValueError("spare BEV channels need the channels_last layout")
        # spare_channels: the BEV is the first C channels of a channels-last [B,X,Y,C+spare] buffer, so
        # that add_target_bev (model/parking_model.py:28-46) fills in its channel without a torch.cat copy
        bev_dt = torch.bfloat16 if shape.bev_dtype == LS_BF16 else torch.float32
        if shape.bev_dtype == LS_BF16 and (bev_format != torch.channels_last or shape.C != 64 or spare_channels):
            raise ValueError("a bf16 BEV tensor needs the dense channels_last layout and 64 channels")
        bev = torch.empty((shape.B, shape.C + spare_channels, shape.X, shape.Y), dtype=bev_dt, device=dev,
                          memory_format=bev_format)
        if spare_channels:
            bev = bev[:, :shape.C]
        prob = torch.empty_like(logits_c)
        st = _bev_strides(bev)
        if rig_cache is not None:
            blob = rig_cache.prepare(shape, code, dev)
            if layout == LS_FEAT_NHWC:
                saved = None                        # the cache carries everything else the backward needs
            check(_lib.load().ls_forward_cached(_ptr(feat_c), layout, _ptr(logits_c), code, _ptr(M.contiguous()),
                                                _ptr(t.contiguous()), _ptr(frustum.contiguous()), C.byref(shape),
                                                _ptr(scratch), scratch.numel(), _ptr(saved),
                                                0 if saved is None else saved.numel(), _ptr(blob), blob.numel(),
                                                int(not rig_cache.valid), _ptr(bev), C.byref(st), _ptr(prob),
                                                _stream(feat)), "ls_forward_cached")
            rig_cache.valid = True
        else:
            check(_lib.load().ls_forward(_ptr(feat_c), layout, _ptr(logits_c), code, _ptr(M.contiguous()),
                                         _ptr(t.contiguous()), _ptr(frustum.contiguous()), C.byref(shape),
                                         _ptr(scratch), scratch.numel(), _ptr(saved),
                                         0 if saved is None else saved.numel(), _ptr(bev), C.byref(st), _ptr(prob),
                                         _stream(feat)), "ls_forward")
        ctx.shape, ctx.code, ctx.layout, ctx.saved_blob = shape, code, layout, saved
        ctx.rig_cache, ctx.need_bwd = rig_cache, need_bwd
        ctx.feat_shape = feat.shape
        ctx.save_for_backward(prob, feat_c if layout == LS_FEAT_NHWC else None)
        return bev, prob

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_bev, grad_prob):
        prob, feat_nhwc = ctx.saved_tensors
        shape, code, layout, saved = ctx.shape, ctx.code, ctx.layout, ctx.saved_blob
        if not ctx.need_bwd:
            raise RuntimeError("lift-splat forward ran without backward state")
        dev = prob.device
        if grad_bev is None:
            grad_bev = torch.zeros(shape.B, shape.C, shape.X, shape.Y, dtype=torch.float32, device=dev)
        if shape.bev_dtype == LS_BF16 and not (grad_bev.dtype == torch.bfloat16 and _grad_layout_ok(grad_bev, shape.D)):
            shape = LsShape.from_buffer_copy(shape)      # this gradient cannot be read as bf16 rows: float32 path
            shape.bev_dtype = LS_F32
        if shape.bev_dtype == LS_F32:
            grad_bev = grad_bev.to(torch.float32)
            if not _grad_layout_ok(grad_bev):
                grad_bev = grad_bev.contiguous()
        if grad_prob is not None:
            grad_prob = grad_prob.to(prob.dtype).contiguous()
        fmt = torch.channels_last if layout == LS_FEAT_NHWC else torch.contiguous_format
        gfeat = torch.empty(ctx.feat_shape, dtype=prob.dtype, device=dev, memory_format=fmt)
        glogits = torch.empty_like(prob)
        scratch = scratch_for(scratch_bytes(shape, code, True), dev)
        st = _bev_strides(grad_bev)
        if ctx.rig_cache is not None:
            blob = ctx.rig_cache.blob
            check(_lib.load().ls_backward_cached(_ptr(grad_bev), C.byref(st), _ptr(grad_prob), _ptr(prob),
                                                 _ptr(feat_nhwc), layout, code, C.byref(shape), _ptr(scratch),
                                                 scratch.numel(), _ptr(saved), 0 if saved is None else saved.numel(),
                                                 _ptr(blob), blob.numel(), _ptr(gfeat), _ptr(glogits), _stream(prob)),
                  "ls_backward_cached")
        else:
            check(_lib.load().ls_backward(_ptr(grad_bev), C.byref(st), _ptr(grad_prob), _ptr(prob), _ptr(feat_nhwc),
                                          layout, code, C.byref(shape), _ptr(scratch), scratch.numel(), _ptr(saved),
                                          saved.numel(), _ptr(gfeat), _ptr(glogits), _stream(prob)), "ls_backward")
        return gfeat, glogits, None, None, None, None, None, None, None


def lift_splat(feat: torch.Tensor, depth_logits: torch.Tensor, M: torch.Tensor, t: torch.Tensor,
               frustum: torch.Tensor, grid: GridSpec, bev_format=torch.contiguous_format, spare_channels: int = 0,
               geom_policy: int = 0, rig_cache: Optional[RigCache] = None, bev_dtype=torch.float32
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """feat [B*N,C,fh,fw], depth_logits [B*N,D,fh,fw] (CamEncoder outputs,
    model/cam_encoder.py:102-111), M [B,N,3,3], t [B,N,3], frustum [D,fh,fw,3] ->
    (bev f32[B,C,X,Y] in ``bev_format``, depth_prob [B*N,D,fh,fw])."""
    B, N = M.shape[:2]
    bn, Cc, fh, fw = feat.shape
    D = depth_logits.shape[1]
    if bn != B * N or depth_logits.shape[0] != bn or tuple(depth_logits.shape[2:]) != (fh, fw):
        raise ValueError("feat %s / depth %s do not match %d x %d cameras" %
                         (tuple(feat.shape), tuple(depth_logits.shape), B, N))
    if tuple(frustum.shape) != (D, fh, fw, 3):
        raise ValueError("frustum %s does not match depth/feature maps" % (tuple(frustum.shape),))
    if feat.dtype != depth_logits.dtype:          # e.g. one head left in float32 under autocast: compute in float32
        feat, depth_logits = feat.float(), depth_logits.float()
    if bev_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("bev_dtype must be torch.float32 or torch.bfloat16")
    shape = make_shape(B, N, D, fh, fw, Cc, grid, geom_policy, pick_tile_x(Cc, bev_format == torch.channels_last),
                       LS_BF16 if bev_dtype == torch.bfloat16 else LS_F32)
    return LiftSplatFunction.apply(feat, depth_logits, M, t, frustum, shape, bev_format, spare_channels, rig_cache)


# --------------------------------------------------------------------------------------
# compatibility: proj_bev_feature(geom, x) on a MATERIALISED frustum tensor
# --------------------------------------------------------------------------------------
class ProjBevFunction(torch.autograd.Function):
    """``BevModel.proj_bev_feature(geom, image_feature)`` of the reference API
    (model/bev_model.py:74-107): voxelise the given ego-frame coordinates, keep / rank / sort,
    VoxelsSumming, scatter - for callers that already hold the B x N x D x h x w x C tensor the
    fused path never builds.  Runs on the same kernels: every frustum point is treated as a
    "pixel" of its own with one depth bin of weight 1, its C-vector as that pixel's feature row
    (so the only extra cost is the contiguous [B*Npts, C] copy of ``image_feature``).
    Gradient: every kept point receives its cell's gradient row (tool/geometry.py:307-317);
    ``geom`` gets none (cut by ``.long()``, bev_model.py:86)."""

    @staticmethod
    def forward(ctx, geom, x, grid: GridSpec, bev_format):
        _need_cuda(geom, x)
        lib = _lib.load()
        B = x.shape[0]
        Cc = x.shape[-1]
        npts = x[0].numel() // Cc
        if geom.shape[0] != B or geom[0].numel() != 3 * npts:
            raise ValueError("geom %s does not match image_feature %s" % (tuple(geom.shape), tuple(x.shape)))
        code = _dtype_code(x)
        cp = padded_channels(Cc)
        rows = x.reshape(B * npts, Cc)
        if cp != Cc:
            rows = torch.nn.functional.pad(rows, (0, cp - Cc))
        rows = rows.contiguous()
        # one "camera", one depth bin, fh x fw = Npts pseudo-pixels (fw: any divisor, for grid shapes)
        fw = next(f for f in (1024, 512, 256, 128, 64, 32, 16, 8, 4, 2, 1) if npts % f == 0)
        shape = make_shape(B, 1, 1, npts // fw, fw, Cc, grid)
        dev = x.device
        g32 = geom.detach().to(torch.float32).reshape(B, npts, 3).contiguous()
        tiles, cells, stride = grid_cells(shape)
        cell = torch.empty(B, npts, dtype=torch.int32, device=dev)
        within = torch.empty_like(cell)
        counts = torch.zeros(B, cells, dtype=torch.int32, device=dev)
        check(lib.ls_index_geom(_ptr(g32), C.byref(shape), None, _ptr(cell), _ptr(within), _ptr(counts), _stream(x)),
              "ls_index_geom")
        ones = torch.ones(B, 1, npts // fw, fw, dtype=x.dtype, device=dev)
        need_bwd = bool(ctx.needs_input_grad[1])
        seg, order, recs, pix = sort(cell, within, counts, ones, shape, with_pixel_index=need_bwd)
        recs2 = torch.empty(B, int(lib.ls_sorted_records(C.byref(shape))), 2, dtype=torch.int32, device=dev)
        bev = torch.empty((B, Cc, shape.X, shape.Y), dtype=torch.float32, device=dev, memory_format=bev_format)
        st = _bev_strides(bev)
        check(lib.ls_splat_fwd(_ptr(rows), code, _ptr(recs), _ptr(seg), _ptr(order), _ptr(recs2), C.byref(shape),
                               _ptr(bev), C.byref(st), _stream(x)), "ls_splat_fwd")
        ctx.shape, ctx.code, ctx.x_shape, ctx.cp = shape, code, x.shape, cp
        if need_bwd:
            ctx.save_for_backward(rows, pix, seg)
        return bev

    @staticmethod
    def backward(ctx, grad_bev):
        rows, pix, seg = ctx.saved_tensors
        shape, code = ctx.shape, ctx.code
        dev = rows.device
        grad_bev = grad_bev.to(torch.float32)
        if not _grad_layout_ok(grad_bev):
            grad_bev = grad_bev.contiguous()
        npts = rows.shape[0] // shape.B
        gT = torch.empty(shape.B, shape.X * shape.Y + 1, ctx.cp, dtype=torch.float32, device=dev)
        gprob = torch.empty(rows.shape[0], dtype=torch.float32, device=dev)
        grows = torch.empty_like(rows)
        st = _bev_strides(grad_bev)
        check(_lib.load().ls_splat_bwd(_ptr(grad_bev), C.byref(st), _ptr(rows), code, _ptr(pix), _ptr(seg),
                                       C.byref(shape), _ptr(gT), _ptr(gprob), _ptr(grows), _stream(rows)),
              "ls_splat_bwd")
        return None, grows[:, :ctx.x_shape[-1]].reshape(ctx.x_shape), None, None


def proj_bev(geom: torch.Tensor, image_feature: torch.Tensor, grid: GridSpec,
             bev_format=torch.contiguous_format) -> torch.Tensor:
    """geom [B,N,D,h,w,3], image_feature [B,N,D,h,w,C] -> bev f32[B,C,X,Y]."""
    return ProjBevFunction.apply(geom, image_feature, grid, bev_format)
