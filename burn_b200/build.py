"""In-tree build of libburn_b200.so (sm_100a only) with plain nvcc.

`python -m burn_b200.build` compiles every .cu under burn_b200/csrc into
burn_b200/lib/libburn_b200.so.  Objects are rebuilt only when a source or header
is newer.  nvcc cross-compiles without a GPU, so this also runs on the CPU-only
build box (driver's `__graft_entry__.build()` check).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
CSRC = ROOT / "csrc"
HOST = ROOT / "host"
OBJ = ROOT / "build"
LIBDIR = ROOT / "lib"
LIB = LIBDIR / "libburn_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = [
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "--expt-extended-lambda",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "-I", str(REPO / "include"), "-I", str(CSRC),
    "-DB200_BUILDING=1",
]


def _sources() -> list[Path]:
    srcs = sorted(CSRC.glob("*.cu"))
    srcs += sorted(HOST.glob("*.cpp"))
    return srcs


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((REPO / "include").glob("*.h"))
    hs += list(HOST.glob("*.h")) + list(HOST.glob("*.hpp"))
    return max((h.stat().st_mtime for h in hs), default=0.0)


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    newest = max(src.stat().st_mtime, _headers_mtime(), Path(__file__).stat().st_mtime)
    if obj.exists() and obj.stat().st_mtime > newest:
        return obj
    cmd = [NVCC, *ARCH, *COMMON, "-c", str(src), "-o", str(obj)]
    if src.suffix == ".cpp":
        cmd.insert(1, "-x")
        cmd.insert(2, "cu")
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed on {src.name}")
    if verbose or r.stderr.strip():
        sys.stderr.write(r.stderr)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest_obj = max(o.stat().st_mtime for o in objs)
    if LIB.exists() and LIB.stat().st_mtime > newest_obj and not force:
        return LIB
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs),
           "-Xcompiler", "-fPIC", "-cudart", "shared", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libburn_b200.so failed")
    return LIB


if __name__ == "__main__":
    out = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(out)
