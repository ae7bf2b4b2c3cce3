"""Builds libssnt_tts_c.so (CUDA kernels + C-ABI) for sm_100a with nvcc, in-tree.

    python ssnt-tts-rust_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssnt_tts_c.so")
SOURCES = ["runtime.cu", "host_copy.cu", "fb_kernels.cu", "tone_fb_kernels.cu", "tone_bf.cu", "tone_ws.cu", "beam_kernels.cu", "trace_kernels.cu",
           "edit_distance.cu", "c_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-fmad=true",
] + (["-DSSNT_BF_DEBUG_VARIANTS"] if os.environ.get("SSNT_BF_DEBUG_VARIANTS") else []) \
  + (["-DSSNT_TP_TRACE"] if os.environ.get("SSNT_TP_TRACE") else []) \
  + (["-DSSNT_TP_TIMING"] if os.environ.get("SSNT_TP_TIMING") else []) \
  + (["-DSSNT_TP_L=" + os.environ["SSNT_TP_L"]] if os.environ.get("SSNT_TP_L") else []) \
  + (["-DSSNT_WS_BWD_STAGES=" + os.environ["SSNT_WS_BWD_STAGES"]] if os.environ.get("SSNT_WS_BWD_STAGES") else []) \
  + (["-DSSNT_WS_FWD_STAGES=" + os.environ["SSNT_WS_FWD_STAGES"]] if os.environ.get("SSNT_WS_FWD_STAGES") else [])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ssnt_tts_c.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-cudart", "static"]
    subprocess.check_call(link)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
