"""In-tree build of the C-ABI library: ``nvcc`` -> ``freegaussian_b200/libfreegaussian_b200.so``.

sm_100a only, ``-lineinfo`` so ncu's source page maps back to the .cu files.  The library
links nothing but the CUDA runtime; it has no torch dependency (include/fg_api.h).
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libfreegaussian_b200.so"
STAMP = PKG / ".libfreegaussian_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O3",
    "-diag-suppress", "177",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "fg_api.h"]):
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the sm_100a kernels cannot be built")
    return exe


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library (object files in build/)."""
    fp = _fingerprint()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == fp:
        return LIB
    objdir = PKG.parent / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-c", str(src), "-o", str(obj)]
        log = open(objdir / (src.stem + ".log"), "w")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log))
    objs = []
    for src, obj, proc, log in procs:
        rc = proc.wait()
        log.close()
        text = (objdir / (src.stem + ".log")).read_text()
        if rc != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{text}")
        if verbose:
            print(text)
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n" + out.stdout + out.stderr)
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
