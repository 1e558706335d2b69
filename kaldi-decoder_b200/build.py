"""In-tree build of the CUDA library and the Python extension (sm_100a only).

    python kaldi-decoder_b200/build.py            # build what is out of date
    python kaldi-decoder_b200/build.py --force

Outputs (git-ignored, shipped to the GPU box with the snapshot):
    kaldi-decoder_b200/lib/libkd_b200.so                     C ABI + kernels
    kaldi-decoder_b200/python/kaldi_decoder/lib/_kaldi_decoder*.so   pybind11 module
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
PY_LIB_DIR = os.path.join(HERE, "python", "kaldi_decoder", "lib")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_cuda(force: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, "libkd_b200.so")
    srcs = [os.path.join(CSRC, "kd_capi.cu"), os.path.join(CSRC, "kd_kernels.cuh"),
            os.path.join(ROOT, "include", "kd_capi.h")]
    if force or _newer(out, srcs):
        _run([NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I" + os.path.join(ROOT, "include"),
              "-I" + CSRC, srcs[0], "-o", out])
    return out


def build_pybind(force: bool = False) -> str:
    import pybind11
    os.makedirs(PY_LIB_DIR, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(PY_LIB_DIR, "_kaldi_decoder" + suffix)
    names = ["faster-decoder.cc", "simple-decoder.cc", "decodable-ctc.cc", "fst-io.cc",
             os.path.join("python", "module.cc")]
    srcs = [os.path.join(CSRC, n) for n in names]
    # every header below csrc/ (minifst/, kaldifst stand-ins, compat shims) and the C ABI
    hdrs = [os.path.join(d, n) for d, _, fs in os.walk(CSRC) for n in fs if n.endswith(".h")]
    hdrs.append(os.path.join(ROOT, "include", "kd_capi.h"))
    if not all(os.path.exists(s) for s in srcs):
        return ""
    if force or _newer(out, srcs + hdrs + [os.path.join(LIB_DIR, "libkd_b200.so")]):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
              "-I" + os.path.join(ROOT, "include"), "-I" + ROOT, "-I" + os.path.join(CSRC, "minifst"),
              "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
              *srcs, "-L" + LIB_DIR, "-lkd_b200", "-Wl,-rpath,$ORIGIN/../../../lib", "-o", out])
    return out


def build_all(force: bool = False):
    a = build_cuda(force)
    b = build_pybind(force)
    return a, b


if __name__ == "__main__":
    print(build_all("--force" in sys.argv))
