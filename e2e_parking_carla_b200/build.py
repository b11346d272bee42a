"""In-tree build of the CUDA library (nvcc, sm_100a only).

``python -m e2e_parking_carla_b200.build`` or ``__graft_entry__.build()``.  The
shared object lands next to this file (git-ignored, but it travels with the
repo snapshot to the GPU box).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libls_b200.so")
SOURCES = ["ls_index.cu", "ls_dense.cu", "ls_splat.cu", "ls_loss.cu", "ls_api.cu"]
HEADERS = ["ls_common.cuh", "ls_internal.h", os.path.join(REPO_ROOT, "include", "ls_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # no --use_fast_math: voxel indices must be bit-exact (IEEE divide, no FMA contraction
    # is enforced in the source with __fmul_rn/__fadd_rn/__fdiv_rn)
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES]
    deps += [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu (one nvcc per translation unit, in parallel) and link
    libls_b200.so; skipped when up to date."""
    if not force and not _stale():
        return LIB_PATH
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    base = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(REPO_ROOT, "include"), "-I", CSRC]
    if verbose:
        base += ["-Xptxas", "-v"]
    for macro in ("LS_TX", "LS_TY", "LS_PDL_TRIGGER", "LS_IDX_ILP", "LS_TCHUNK", "LS_QWIN", "LS_SPLAT_MINB", "LS_PLACE_GROUPS", "LS_GATHER_THREADS", "LS_GATHER_MINB", "LS_GATHER_REVERSE", "LS_GATHER_OCC", "LS_GOCC_ROWS", "LS_GOCC_MINB", "LS_SPLATD_MINB", "LS_CANON_BIG", "LS_GATHER_SKIP_DEAD", "LS_FFMA2", "LS_GL_DIRECT", "LS_SPLAT_FASTLOOP", "LS_CANON_THREADS", "LS_GATHER_L2PF", "LS_EPI_THREADS", "LS_IDX_PIPE", "LS_ABLATE", "LS_GOCC_BFLY8", "LS_GATHER_PF1"):     # developer knobs: tile shape, early-launch trigger
        if os.environ.get(macro):
            base += ["-D%s=%s" % (macro, os.environ[macro])]
    if os.environ.get("LS_PROFILE"):
        base += ["-DLS_PROFILE"]      # developer build: per-phase clock64 accounting in the splat
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = base + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    objs = []
    for cmd, obj, pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), err))
        if verbose:
            sys.stderr.write(err)
        objs.append(obj)
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(link), res.stderr))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
