"""Builds the C-ABI shared library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dualdiffusion_b200.build [--force]

Output: dualdiffusion_b200/lib/libdualdiffusion_b200.so (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdualdiffusion_b200.so")
STAMP_PATH = os.path.join(LIB_DIR, "build.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inl")))
    files.append(os.path.join(ROOT, "include", "dualdiffusion_b200.h"))
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())      # relative: the same tree fingerprints equally wherever it is copied
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _up_to_date(fp: str) -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == fp


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the sources changed since the library was built.  Safe under torchrun: one process compiles (file lock),
    the others wait and find the stamp; the library appears atomically."""
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    fp = _fingerprint()
    if not force and _up_to_date(fp):
        return LIB_PATH
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and _up_to_date(fp):          # another process built it while we waited
            return LIB_PATH
        return _build_locked(fp, verbose)


def _build_locked(fp: str, verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(LIB_DIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see dualdiffusion_b200/lib/build.log")
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", tmp, *objs, "-Xlinker", "--exclude-libs", "-Xlinker", "ALL"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB_PATH)
    with open(STAMP_PATH, "w") as fh:
        fh.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
