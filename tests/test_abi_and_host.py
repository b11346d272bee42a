"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every
symbol declared in include/ls_b200.h (no compute calls without a GPU), argument checking
works across the ABI, and the host mirror of BevModel keeps the reference's contract."""
import ctypes as C
import os
import re

import pytest
import torch

from e2e_parking_carla_b200.synthetic import LiftSplatShape, make_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ls_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ls_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from e2e_parking_carla_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libls_b200.so does not export %s" % name
    # and the ctypes prototypes cover exactly the header
    assert sorted(_lib.PROTOTYPES) == declared
    assert b"sm_100a" in lib.ls_version()


def test_library_targets_sm100a_only():
    """The shipped cubin is sm_100a: no multi-arch fatbin, no PTX for other vendors/archs."""
    import subprocess
    from e2e_parking_carla_b200.build import LIB_PATH, build
    build()
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_abi_argument_checking(lib):
    """Bad arguments come back as status codes (never exceptions / crashes across the ABI)."""
    from e2e_parking_carla_b200 import _lib
    from e2e_parking_carla_b200 import lift_splat as ls
    grid = ls.GridSpec((-9.95, -9.95, 0.0), (0.1, 0.1, 20.0), (200, 200, 1))
    s = ls.make_shape(1, 4, 48, 32, 32, 64, grid)
    dummy16 = C.c_void_p(16)
    tiles, cells, stride = ls.grid_cells(s)
    tile_cells = cells // tiles
    assert tile_cells in (128, 256)
    assert tiles * tile_cells == cells >= 200 * 200 and stride == cells + 4
    assert ls.padded_channels(6) == 8 and ls.padded_channels(64) == 64
    assert ls.scratch_bytes(s, ls.LS_F32, True) > ls.scratch_bytes(s, ls.LS_F32, False) > 0
    # backward state: a channels_last feature map is consumed in place, so less is saved
    assert ls.saved_bytes(s, ls.LS_F32, ls.LS_FEAT_NCHW) > ls.saved_bytes(s, ls.LS_F32, ls.LS_FEAT_NHWC) > 0
    # null pointers
    assert lib.ls_camera_transform(None, None, 4, None, None, None) == -1
    assert lib.ls_index(None, None, None, C.byref(s), None, None, None, None, None) == -1
    # Z != 1 is unsupported for the splat (reference squeezes Z, model/bev_model.py:104)
    bad = ls.make_shape(1, 4, 48, 32, 32, 64, ls.GridSpec(grid.start, grid.res, (200, 200, 2)))
    assert lib.ls_scratch_bytes(C.byref(bad), ls.LS_F32, 1) == 0
    with pytest.raises(ValueError):
        ls.scratch_bytes(bad, ls.LS_F32, True)
    # limits of the placement kernel are reported up front, not in the middle of ls_forward
    deep = ls.make_shape(1, 4, 192, 32, 32, 64, grid)
    assert lib.ls_scratch_bytes(C.byref(deep), ls.LS_F32, 1) == 0
    assert lib.ls_scratch_bytes(C.byref(ls.make_shape(1, 4, 191, 8, 8, 64, grid)), ls.LS_F32, 1) > 0
    # channels_last features need whole 16-byte channel quads
    assert lib.ls_saved_bytes(C.byref(ls.make_shape(1, 4, 48, 32, 32, 6, grid)), ls.LS_F32, ls.LS_FEAT_NHWC) == 0
    # BEV strides: neither NCHW-like (y == 1) nor channels-last-like (c == 1) is refused
    odd = ls.LsBevStrides(64 * 200 * 200, 2, 200 * 128, 128)
    assert lib.ls_forward(dummy16, 0, dummy16, 0, dummy16, dummy16, dummy16, C.byref(s), dummy16, 1 << 40, None, 0,
                          dummy16, C.byref(odd), dummy16, None) == -2
    # round-2 entry points: same conventions
    assert lib.ls_cache_bytes(C.byref(s)) > 0 and lib.ls_cache_bytes(C.byref(bad)) == 0
    assert lib.ls_forward_cached(None, 0, None, 0, None, None, None, C.byref(s), None, 0, None, 0, None, 0, 1, None,
                                 None, None, None) == -1
    assert lib.ls_forward_cached(dummy16, 0, dummy16, 0, dummy16, dummy16, dummy16, C.byref(s), dummy16, 1 << 40, None, 0,
                                 dummy16, 16, 1, dummy16, C.byref(ls.LsBevStrides(2560000, 40000, 200, 1)), dummy16,
                                 None) == -3                       # cache blob too small
    assert lib.ls_depth_loss_ws_bytes(4, 32, 32) == 4 * 32 * 2 * 4 and lib.ls_depth_loss_ws_bytes(0, 32, 32) == 0
    assert lib.ls_depth_loss_fwd(None, 0, None, 4, 48, 32, 32, 8, 0.25, 0.25, None, None, 0, None, None) == -1
    assert lib.ls_depth_loss_fwd(dummy16, 0, dummy16, 4, 48, 32, 32, 8, 0.25, 0.25, dummy16, dummy16, 8, dummy16,
                                 None) == -3                       # workspace too small
    assert lib.ls_depth_loss_bwd(dummy16, 3, dummy16, dummy16, None, 4, 48, 32, 32, dummy16, None) == -1   # dtype
    # launch limits (grid.y = B*N, 32-bit pixel and block counts) are reported, not wrapped
    assert lib.ls_depth_loss_ws_bytes(65535, 32, 32) > 0 and lib.ls_depth_loss_ws_bytes(65536, 32, 32) == 0
    assert lib.ls_depth_loss_ws_bytes(4, 65535, 65535) == 0 and lib.ls_depth_loss_ws_bytes(65535, 4096, 4096) == 0
    assert lib.ls_depth_loss_fwd(dummy16, 0, dummy16, 70000, 48, 32, 32, 8, 0.25, 0.25, dummy16, dummy16, 1 << 40,
                                 dummy16, None) == -2
    assert lib.ls_depth_loss_bwd(dummy16, 0, dummy16, dummy16, None, 4, 48, 65535, 65535, dummy16, None) == -2
    assert lib.ls_target_bev(None, 2, 200, 200, None, 0, 0, 0, None) == -1
    assert lib.ls_index_geom(None, C.byref(s), None, None, None, None, None) == -1
    odd_policy = ls.make_shape(1, 4, 48, 32, 32, 64, grid, geom_policy=7)
    assert lib.ls_scratch_bytes(C.byref(odd_policy), ls.LS_F32, 1) == 0
    assert lib.ls_scratch_bytes(C.byref(ls.make_shape(1, 4, 48, 32, 32, 64, grid, tile_x=3)), ls.LS_F32, 1) == 0
    assert lib.ls_scratch_bytes(C.byref(ls.make_shape(1, 4, 48, 32, 32, 64, grid, tile_x=8)), ls.LS_F32, 1) > 0
    # unknown dtype code
    dummy = C.c_void_p(16)
    assert lib.ls_softmax(dummy, 7, C.byref(s), dummy, None) == -1
    assert lib.ls_strerror(-2) == b"unsupported configuration"
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(_lib.LiftSplatLibraryError):
        _lib.check(-3, "x")


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    from e2e_parking_carla_b200 import BevModel
    from e2e_parking_carla_b200 import lift_splat as ls
    with pytest.raises(RuntimeError, match="CUDA"):
        ls.camera_transform(torch.eye(3).expand(1, 4, 3, 3), torch.eye(4).expand(1, 4, 4, 4))
    model = BevModel(make_cfg(LiftSplatShape()), cam_encoder=torch.nn.Identity())
    with pytest.raises(RuntimeError, match="CUDA"):
        model.get_geometry(torch.eye(3).expand(1, 4, 3, 3), torch.eye(4).expand(1, 4, 4, 4))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "e2e_parking_carla_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_bev_model_keeps_reference_contract():
    """Constructor argument, attributes, state_dict entries and dtypes of the reference's
    BevModel (model/bev_model.py:10-26)."""
    from e2e_parking_carla_b200 import BevModel, calculate_birds_eye_view_parameters

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(3))

    cfg = make_cfg(LiftSplatShape())
    m = BevModel(cfg, cam_encoder=Enc())
    sd = m.state_dict()
    assert list(sd) == ["bev_res", "bev_start_pos", "bev_dim", "frustum", "cam_encoder.w"]
    assert sd["bev_dim"].dtype == torch.int64 and sd["bev_dim"].tolist() == [200, 200, 1]
    assert sd["bev_res"].dtype == torch.float32 and sd["frustum"].shape == (48, 32, 32, 3)
    assert not any(p.requires_grad for n, p in m.named_parameters() if not n.startswith("cam_encoder"))
    assert m.depth_channel == 48 and m.down_sample == 8 and m.cfg is cfg
    for name in ("create_frustum", "get_geometry", "encoder_forward", "proj_bev_feature", "calc_bev_feature", "forward"):
        assert callable(getattr(m, name))
    res, start, dim = calculate_birds_eye_view_parameters([-10.0, 10.0, 0.05], [-10.0, 10.0, 0.05], [-10.0, 10.0, 20.0])
    assert dim.tolist() == [400, 400, 1] and dim.dtype == torch.int64
    # strict load of a reference-shaped checkpoint (agent/parking_agent.py:260-262)
    m2 = BevModel(cfg, cam_encoder=Enc())
    m2.load_state_dict(sd, strict=True)
    # configurations the reference cannot run either are rejected up front
    bad = make_cfg(LiftSplatShape())
    bad.use_depth_distribution = 0
    with pytest.raises(ValueError):
        BevModel(bad, cam_encoder=Enc())
    bad = make_cfg(LiftSplatShape(bev_z_bound=[-10.0, 10.0, 10.0]))
    with pytest.raises(ValueError):
        BevModel(bad, cam_encoder=Enc())


def test_reference_state_dict_loads(tmp_path):
    """A state_dict produced by the reference's own BevModel loads strictly (when the
    reference tree is available in this container)."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    from e2e_parking_carla_b200 import BevModel
    cfg = make_cfg(LiftSplatShape())
    ref = rh.reference_bev_model(cfg)
    ours = BevModel(cfg, cam_encoder=torch.nn.Identity())
    sd = {k: v for k, v in ref.state_dict().items() if not k.startswith("cam_encoder")}
    ours.load_state_dict(sd, strict=True)
    for k, v in sd.items():
        assert torch.equal(getattr(ours, k).data, v) and getattr(ours, k).dtype == v.dtype


def test_integration_doc_stub_matches_the_header():
    """The ctypes stub INTEGRATION.md shows a maintainer declares LsShape / LsBevStrides with the
    fields of include/ls_b200.h in order (a stale stub would hand the library a short struct), and
    every entry point its table names is declared in the header."""
    from e2e_parking_carla_b200 import _lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = doc[doc.index("class LsShape(C.Structure)"):doc.index("class LsBevStrides(C.Structure)")]
    ints = re.search(r'for n in \(([^)]*)\)', stub).group(1)
    names = re.findall(r'"(\w+)"', ints) + re.findall(r'\("(\w+)", C\.c_', stub)
    assert names == [f[0] for f in _lib.LsShape._fields_]
    strides = doc[doc.index("class LsBevStrides(C.Structure)"):]
    assert re.findall(r'"(\w+)"', re.search(r'for n in \(([^)]*)\)', strides).group(1)) == \
        [f[0] for f in _lib.LsBevStrides._fields_]
    table = doc[doc.index("## Entry points and the reference lines they replace"):doc.index("## Streams, graphs")]
    declared = set(_declared_symbols())
    named = set(re.findall(r"`(ls_[a-z0-9_]+)`", table))
    assert named and named <= declared, named - declared
    # the table covers the whole ABI except the introspection helpers
    helpers = {"ls_version", "ls_strerror", "ls_last_cuda_error", "ls_launch_count", "ls_debug_phase_cycles",
               "ls_grid_cells", "ls_padded_channels", "ls_sorted_records"}
    assert declared - named <= helpers, declared - named - helpers


def test_size_functions_never_accept_a_wrapped_shape(lib):
    """Seeded fuzz of the shape checker behind ls_scratch_bytes / ls_saved_bytes / ls_cache_bytes (all
    host-only): whatever it accepts obeys the documented limits in exact integer arithmetic (a 64-bit
    product of five int32 sizes can wrap; a wrapped product must not pass), the three sizes are
    positive, far from wrapping, agree on acceptance and grow with the batch."""
    import random
    from e2e_parking_carla_b200 import lift_splat as ls
    rnd = random.Random(7)
    vals = [0, 1, 2, 3, 4, 7, 8, 16, 31, 32, 33, 48, 63, 64, 65, 96, 128, 191, 192, 200, 255, 256, 257, 400, 1000, 1024,
            4096, 65535, 65536, 1 << 20, (1 << 31) - 1, -1, -5]
    accepted = 0
    for _ in range(30000):
        B, X, Y = rnd.choice(vals), rnd.choice(vals), rnd.choice(vals)
        N, D, Cc = rnd.choice(vals[:20] + [-1]), rnd.choice(vals[:24]), rnd.choice(vals[:24])
        fh, fw = rnd.choice(vals[:28] + [65535]), rnd.choice(vals[:28] + [65535])
        Z, tx = rnd.choice([1, 1, 1, 0, 2]), rnd.choice([0, 1, 2, 4, 8, 16, 32, 64, 128, 3, 256, -1])
        grid = ls.GridSpec((-9.95, -9.95, 0.0), (0.1, 0.1, 20.0), (X, Y, Z))
        s = ls.make_shape(B, N, D, fh, fw, Cc, grid, 0, tx)
        for code in (ls.LS_F32, ls.LS_BF16):
            full = lib.ls_scratch_bytes(C.byref(s), code, 1)
            fwd_only = lib.ls_scratch_bytes(C.byref(s), code, 0)
            saved = lib.ls_saved_bytes(C.byref(s), code, ls.LS_FEAT_NCHW)
            cache = lib.ls_cache_bytes(C.byref(s))
            assert (full > 0) == (fwd_only > 0) == (saved > 0) == (cache > 0), (B, N, D, fh, fw, Cc, X, Y, Z, tx)
            if not full:
                continue
            accepted += 1
            assert min(B, N, D, fh, fw, Cc, X, Y) >= 1 and Z == 1
            assert B * N * D * fh * fw < 1 << 31 and X * Y * Z < 1 << 28 and Cc <= 256 and D <= 191
            assert N * fh * fw <= 1 << 20
            assert fwd_only <= full < 1 << 60 and saved < 1 << 60 and cache < 1 << 60
            if B > 1:
                half = ls.make_shape(B // 2, N, D, fh, fw, Cc, grid, 0, tx)
                assert 0 < lib.ls_scratch_bytes(C.byref(half), code, 1) <= full
                assert 0 < lib.ls_saved_bytes(C.byref(half), code, ls.LS_FEAT_NCHW) <= saved
    assert accepted > 1000
    # the shapes that used to wrap: B * N * D * fh * fw = 2.4e21 and 2.5e22
    for B, N, D, fh, fw, Cc, X, Y, tx in ((2147483647, 33, 8, 65535, 65535, 63, 256, 255, 32),
                                          (2147483647, 96, 31, 65535, 65535, 191, 33, 1000, 32)):
        s = ls.make_shape(B, N, D, fh, fw, Cc, ls.GridSpec((0.0, 0.0, 0.0), (0.1, 0.1, 20.0), (X, Y, 1)), 0, tx)
        assert lib.ls_scratch_bytes(C.byref(s), ls.LS_F32, 1) == 0 and lib.ls_cache_bytes(C.byref(s)) == 0


def test_every_entry_point_survives_all_null_arguments(lib):
    """Each of the header's entry points called with NULL / zero for every argument returns
    LS_ERR_BAD_ARG (int status), 0 (a size) or a string - never a crash, never LS_OK - and a valid
    shape with null buffers is refused the same way."""
    from e2e_parking_carla_b200 import _lib
    from e2e_parking_carla_b200 import lift_splat as ls
    good = ls.make_shape(1, 4, 48, 32, 32, 64, ls.GridSpec((-9.95, -9.95, 0.0), (0.1, 0.1, 20.0), (200, 200, 1)))
    for with_shape in (False, True):
        for name, (res, args) in sorted(_lib.PROTOTYPES.items()):
            vals = []
            for a in args:
                if a is _lib._SH and with_shape:
                    vals.append(C.byref(good))
                elif a in (C.c_float, C.c_double):
                    vals.append(0.0)
                elif a is C.c_void_p or getattr(a, "__name__", "").startswith("LP_"):
                    vals.append(None)
                else:
                    vals.append(0)
            out = getattr(lib, name)(*vals)
            if res is C.c_int and name not in ("ls_padded_channels",):
                if with_shape and name == "ls_grid_cells":
                    assert out == 0                        # its three outputs are optional
                else:
                    assert out == -1, (name, out)
            elif res is C.c_char_p:
                assert isinstance(out, bytes)
            elif not with_shape or _lib._SH not in args:
                assert out == 0, (name, out)
    assert lib.ls_launch_count() == 0                      # nothing was launched by any of it
