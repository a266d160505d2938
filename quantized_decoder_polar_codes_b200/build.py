"""In-tree build of the two native artefacts (they travel to the GPU box with the snapshot):

  libpolar_b200.so                  nvcc, sm_100a only   -- the C ABI of include/polar_b200.h + all kernels
  _libPolarDecoder<ext>.so          g++ / pybind11       -- host mirror of the reference's pybind11 module

Usage: python -m quantized_decoder_polar_codes_b200.build [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libpolar_b200.so")
EXT = os.path.join(HERE, "_libPolarDecoder" + sysconfig.get_config_var("EXT_SUFFIX"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false",            # fp64 families: no contraction
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(exts):
    out = [os.path.join(ROOT, "include", "polar_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(exts):
            out.append(os.path.join(CSRC, f))
    return out


def build_cuda(force=False, verbose=False):
    srcs = _sources((".cu", ".cuh", ".h"))
    if force or _newer(LIB, srcs):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, "pb_capi.cu"), "-o", LIB]
        subprocess.check_call(cmd)
    return LIB


def build_pybind(force=False):
    src = os.path.join(CSRC, "pb_pybind.cpp")
    if force or _newer(EXT, [src, os.path.join(ROOT, "include", "polar_b200.h"), LIB]):
        import pybind11
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
               "-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(), src,
               "-L" + HERE, "-lpolar_b200", "-Wl,-rpath,$ORIGIN", "-o", EXT]
        subprocess.check_call(cmd)
    return EXT


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_pybind(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
