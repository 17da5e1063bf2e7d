"""Builds libadn.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent          # audio-denoiser-onnx_b200/
REPO_ROOT = PKG_ROOT.parent
CSRC = PKG_ROOT / "csrc"
LIB_PATH = PKG_ROOT / "libadn.so"
STAMP = PKG_ROOT / ".libadn.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [REPO_ROOT / "include" / "adn.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library.  Rebuilds only when a source,
    header or flag changed (content hash), so the prebuilt .so that travels to the GPU box
    is reused there as-is."""
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    objs = []
    build_dir = PKG_ROOT / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc_path(), *NVCC_FLAGS, "-I", str(REPO_ROOT / "include"), "-I", str(CSRC), "-c", str(src),
               "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            print(f"nvcc failed for {src.name}:\n{out}", file=sys.stderr)
        elif verbose or out.strip():
            print(out, file=sys.stderr)
    if failed:
        raise RuntimeError("libadn build failed")
    link = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
            *map(str, objs), "-lcudart", "-lcuda"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("libadn link failed:\n" + r.stdout)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
