"""Builds libmafb200.so (sm_100a only) in-tree with nvcc.

`python -m maf_yolo_b200.build` or `build_library()`; `__graft_entry__.build()` calls this.
nvcc cross-compiles without a GPU.  The arch is passed as an explicit -gencode pair: the
`-arch=sm_100a` shorthand also emits a plain compute_100 PTX pass, which rejects tcgen05.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
# MAFB200_BUILD_SUFFIX=<name> (with MAFB200_NVCC_EXTRA=-D...) builds an A/B variant next to the default library:
# libmafb200_<name>.so, loaded with MAFB200_LIB=<path> (maf_yolo_b200/_lib.py).  Same sources, same CUDA path.
SUFFIX = os.environ.get("MAFB200_BUILD_SUFFIX", "")
LIB_PATH = PKG_DIR / (f"libmafb200_{SUFFIX}.so" if SUFFIX else "libmafb200.so")
SOURCES = ["host.cu", "gemm_tc.cu", "stem_conv.cu", "dwconv.cu", "dwpw.cu", "bneck.cu", "poolpw.cu", "pool.cu", "decode.cu", "nms.cu", "postprocess.cu", "preprocess.cu", "loss.cu"]
EXTRA = os.environ.get("MAFB200_NVCC_EXTRA", "").split()
NVCC_FLAGS = EXTRA + [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    obj_dir = PKG_DIR / "build" / SUFFIX if SUFFIX else PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True, parents=True)
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "mafb200.h"]
    jobs = []
    objs = []
    for src in SOURCES:
        s = CSRC / src
        o = obj_dir / (src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB_PATH, objs):
        run([nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC", "-lcudart_static", "-ldl", "-lpthread", "-lrt"])
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
