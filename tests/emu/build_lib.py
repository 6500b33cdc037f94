"""Builds tests/emu/_build/libns3d_emu.so: the WHOLE of libns3d.so -- kernels and host code, the
same translation units -- compiled by g++ for the CPU (test infrastructure).

The only change made to the sources is mechanical and happens on a copy: every
``kernel<<<grid, block, smem, stream>>>(args)`` becomes
``emu::launch_on(grid, block, stream, <kernel synchronises?>, [=]() { kernel(args); })``.
``cuda_runtime.h`` and ``nccl.h`` resolve to tests/emu/fake_cuda/.  The result exports the C ABI of
include/ns3d.h, so the test suite can drive it through the ordinary ctypes binding
(``tests.emu.emulated_library()``); nothing under navierstokes3d_b200/ can reach it.
"""
from __future__ import annotations

import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "navierstokes3d_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libns3d_emu.so")
SOURCES = ["ns3d_core.cu", "ns3d_ops.cu", "ns3d_pt.cu", "ns3d_step.cu", "ns3d_ptv.cu", "ns3d_ptv_mode0.cu", "ns3d_ptv_mode1.cu", "ns3d_ptv_mode2.cu", "ns3d_out.cu"]
# kernels that synchronise (__syncthreads, named barriers, warp shuffles): one host thread per CUDA thread
THREADED = ("ptv_kernel_fn", "ptv_flow_kernel_fn", "predictor_kernel", "ptv_residual_kernel", "pt_tb2_kernel", "pt_tb2s_kernel", "pt_tb2sp_kernel", "pt_tb2d_kernel", "pt_residual_kernel", "max_abs_kernel")


def _match_back_template(s: str, end: int) -> int:
    """s[end-1] == '>': index of the matching '<'."""
    depth = 0
    for i in range(end - 1, -1, -1):
        if s[i] == ">":
            depth += 1
        elif s[i] == "<":
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced template arguments")


def _match_paren(s: str, start: int) -> int:
    """s[start] == '(': index just past the matching ')'."""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced parentheses")


def _split_top(s: str) -> list[str]:
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def rewrite_launches(src: str) -> str:
    # the one macro that hides a launch configuration: expand it textually first
    m = re.search(r"#define TBS_ARGS (<<<.*?>>>\(.*?\))\n", src)
    if m:
        body = m.group(1)
        src = src.replace(m.group(0), "")
        src = src.replace("#undef TBS_ARGS\n", "")
        src = re.sub(r"\s+TBS_ARGS\b", lambda _: body, src)
    out, pos = "", 0
    while True:
        a = src.find("<<<", pos)
        if a < 0:
            return out + src[pos:]
        b = src.index(">>>", a)
        # kernel expression: identifier, optionally followed by template arguments, right before '<<<'
        k_end = a
        while src[k_end - 1].isspace():
            k_end -= 1
        k_start = k_end
        if src[k_end - 1] == ">":
            k_start = _match_back_template(src, k_end)
        while k_start > 0 and (src[k_start - 1].isalnum() or src[k_start - 1] == "_"):
            k_start -= 1
        kernel = src[k_start:k_end]
        cfg = _split_top(src[a + 3:b])
        assert len(cfg) == 4, (kernel, cfg)
        p0 = b + 3
        while src[p0].isspace():
            p0 += 1
        assert src[p0] == "(", (kernel, src[p0:p0 + 20])
        p1 = _match_paren(src, p0)
        name = re.match(r"\w+", kernel).group(0)
        threaded = name in THREADED or (name == "pt_iter_kernel" and re.search(r",\s*true\s*>$", kernel) is not None)
        out += src[pos:k_start]
        out += (f"emu::launch_on(({cfg[0]}), ({cfg[1]}), ({cfg[3]}), {'true' if threaded else 'false'}, "
                f"[=]() {{ {kernel}{src[p0:p1]}; }})")
        pos = p1


def build(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps += [os.path.join(HERE, "cuda_host_shim.h"), os.path.join(HERE, "fake_cuda", "cuda_runtime.h"),
             os.path.join(HERE, "fake_cuda", "nccl.h"), os.path.join(HERE, "fake_nccl.cpp"),
             os.path.join(ROOT, "include", "ns3d.h"), __file__]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    objs = []
    procs = []
    for name in SOURCES:
        cpp = os.path.join(OUT, name.replace(".cu", "_emu.cpp"))
        with open(os.path.join(CSRC, name)) as fh:
            text = fh.read()
        # headers that launch kernels themselves (ns3d_ptv_launch.cuh) are inlined so that their launches are rewritten too
        for inc in re.findall(r'#include "(\w+\.cuh)"', text):
            with open(os.path.join(CSRC, inc)) as fh:
                body = fh.read()
            if "<<<" in body:
                text = text.replace(f'#include "{inc}"', body.replace("#pragma once", ""))
        text = rewrite_launches(text)
        with open(cpp, "w") as fh:
            fh.write(f'#line 1 "{os.path.join(CSRC, name)}"\n' + text)
        obj = cpp[:-4] + ".o"
        objs.append(obj)
        cmd = ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-fPIC", "-pthread", "-fvisibility=hidden", "-w",
               "-I", os.path.join(HERE, "fake_cuda"), "-I", CSRC, "-c", cpp, "-o", obj]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)))
    errs = []
    for name, p in procs:
        out, err = p.communicate()
        if p.returncode != 0:
            errs.append(f"--- {name}\n{err[-6000:]}")
    if errs:
        raise RuntimeError("g++ failed building the emulated library:\n" + "\n".join(errs))
    res = subprocess.run(["g++", "-shared", "-pthread", *objs, "-o", LIB, "-ldl", "-lrt"], capture_output=True, text=True,
                         env=env)
    if res.returncode != 0:
        raise RuntimeError("linking the emulated library failed:\n" + res.stderr[-4000:])
    # the fake NCCL the emulated library dlopens in multi-rank runs (LD_LIBRARY_PATH = this directory)
    res = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-pthread",
                          os.path.join(HERE, "fake_nccl.cpp"), "-o", os.path.join(OUT, "libnccl.so.2"), "-lrt"],
                         capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("building the fake NCCL failed:\n" + res.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
