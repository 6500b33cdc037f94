"""Host emulation of the CUDA kernels of the fused PT loop (test infrastructure).

``tests/emu/pt_emu.cpp`` compiles ``navierstokes3d_b200/csrc/ns3d_pt_kernels.cuh`` -- the very
source nvcc compiles into libns3d.so -- with g++ behind ``cuda_host_shim.h`` (CUDA threads = host
threads, ``__syncthreads`` = barrier).  It lets the CPU test suite execute the kernels' index
arithmetic, boundary folding, tiling and ping-pong logic bit for bit against the oracle without
a GPU.  It is NOT a product path (nothing in navierstokes3d_b200/ can reach it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libpt_emu.so")
DEPS = [os.path.join(HERE, "pt_emu.cpp"), os.path.join(HERE, "cuda_host_shim.h"),
        os.path.join(ROOT, "navierstokes3d_b200", "csrc", "ns3d_pt_kernels.cuh"),
        os.path.join(ROOT, "navierstokes3d_b200", "csrc", "ns3d_shared.cuh"),
        os.path.join(ROOT, "include", "ns3d.h")]
_lib = None

KERNELS = {"pt_iter": 0, "pt_tb2": 1, "pt_tb2s": 2}


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    # -ffp-contract=off: like nvcc --fmad=false, FMA only where fma() is written
    cmd = ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread",
           os.path.join(HERE, "pt_emu.cpp"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("g++ failed building the kernel emulation:\n" + res.stderr)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_pt_iterate.restype = C.c_int
        _lib.emu_pt_iterate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
        _lib.emu_pt_tb2_split.restype = C.c_int
        _lib.emu_pt_tb2_split.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int, C.c_int]
    return _lib


def pt_tb2_split(kernel_mid: str, mode: int, pt_params, Pr, dP, divV, n_pairs: int, klo: int, khi: int,
                 ty_mid: int = 16, zchunk_mid: int = 16) -> None:
    """n_pairs double iterations, each as three launches over the plane ranges [1,klo), [klo,khi),
    [khi,nz-1) -- the outer two with pt_tb2_kernel, the middle one with `kernel_mid` (the way slabs
    split every launch into interface chunks and the rest)."""
    rc = lib().emu_pt_tb2_split(KERNELS[kernel_mid], mode, ty_mid, C.addressof(pt_params), Pr.ctypes.data,
                                dP.ctypes.data, divV.ctypes.data, n_pairs, klo, khi, zchunk_mid)
    if rc != 0:
        raise RuntimeError(f"emu_pt_tb2_split failed ({rc})")


def pt_iterate(kernel: str, mode: int, pt_params, Pr: np.ndarray, dP: np.ndarray, divV: np.ndarray, n: int,
               ty: int = 16, serpentine: bool = True, zlo_halo: bool = False, zhi_halo: bool = False) -> int:
    """n fused PT iterations on host arrays (Fortran order, updated in place) through the emulated
    kernels, driven like one rank's run_direct() in ns3d_pt.cu.  Returns the number of launches."""
    for a in (Pr, dP, divV):
        assert a.dtype == np.float64 and a.flags.f_contiguous
    nl = C.c_longlong(0)
    rc = lib().emu_pt_iterate(KERNELS[kernel], mode, ty, C.addressof(pt_params), int(zlo_halo), int(zhi_halo),
                              int(serpentine), Pr.ctypes.data, dP.ctypes.data, divV.ctypes.data, n, C.byref(nl))
    if rc != 0:
        raise RuntimeError(f"emu_pt_iterate failed ({rc})")
    return nl.value
