"""In-tree build of the C-ABI CUDA library (sm_100a only).

`python -m stableavatar_b200.build` compiles every `csrc/*.cu` with nvcc into
`stableavatar_b200/lib/libsa_b200.so`. Objects are rebuilt only when a source or header changed.
The library has no torch dependency: entry points are `extern "C"` (see include/stableavatar_b200.h).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = LIBDIR / "obj"
LIB = LIBDIR / "libsa_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: Path, stamp: str, verbose: bool) -> Path:
    obj = OBJDIR / (src.stem + ".o")
    tag = OBJDIR / (src.stem + ".stamp")
    if obj.exists() and tag.exists() and tag.read_text() == stamp:
        return obj
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = OBJDIR / (src.stem + ".ptxas.log")
    log.write_text(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        for line in r.stderr.splitlines():
            if "registers" in line or "spill" in line or "warning" in line.lower():
                print(f"[{src.name}] {line.strip()}")
    tag.write_text(stamp)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJDIR.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((ROOT / "include").glob("*.h"))
    hstamp = _digest(hdrs)
    if force:
        for f in OBJDIR.glob("*.stamp"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hstamp + _digest([s]), verbose), srcs))
    link_stamp = _digest(objs)
    tag = OBJDIR / "link.stamp"
    if not (LIB.exists() and tag.exists() and tag.read_text() == link_stamp):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        tag.write_text(link_stamp)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
