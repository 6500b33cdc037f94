"""In-tree build of libns3d.so (hand-written CUDA for sm_100a + the C ABI of include/ns3d.h).

``python -m navierstokes3d_b200.build`` or ``build()``; nvcc cross-compiles without a GPU.
The shared object lands next to the sources (navierstokes3d_b200/csrc/libns3d.so) so that it
travels with the repo snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libns3d.so")
SOURCES = ["ns3d_core.cu", "ns3d_ops.cu", "ns3d_pt.cu", "ns3d_step.cu", "ns3d_ptv.cu", "ns3d_ptv_mode0.cu", "ns3d_ptv_mode1.cu", "ns3d_ptv_mode2.cu", "ns3d_out.cu"]
HEADERS = ["ns3d_internal.cuh", "ns3d_shared.cuh", "ns3d_pt_common.cuh", "ns3d_ptv_kernels.cuh", "ns3d_ptv_launch.cuh", os.path.join("..", "..", "include", "ns3d.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # The reference's CPU backend never contracts a*b+c; FMA is used only via explicit fma().
    "--fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libns3d.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, s)) > t for s in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds the library unless it is up to date.  Several ranks of one job may get here at once (torchrun on a
    box whose snapshot has sources newer than the library): one of them builds, the others wait on a file lock."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB   # another process built it while this one waited
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports a gcc wrapper that nvcc must not pick up
    if os.path.exists(LIB):
        os.remove(LIB)  # never leave a stale library behind a failed build
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src: str):
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc, *compile_flags, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", os.path.join(CSRC, src), "-o", obj]
        return obj, subprocess.run(cmd, capture_output=True, text=True, env=env)

    # one nvcc per translation unit, in parallel (ns3d_pt.cu with its kernel instantiations dominates)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = False
    for _, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libns3d.so (see stderr)")
    link = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC", *[o for o, _ in results], "-o", LIB, "-ldl"],
                          capture_output=True, text=True, env=env)
    for obj, _ in results:
        if os.path.exists(obj):
            os.remove(obj)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("nvcc failed linking libns3d.so (see stderr)")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
