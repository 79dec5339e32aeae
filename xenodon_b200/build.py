"""Build recipe: nvcc (sm_100a) + g++ -> xenodon_b200/libxenodon_b200.so and bin/xenodon.

In-tree build so the artefacts travel with the repository snapshot to the GPU box.
`python -m xenodon_b200.build [--force]`
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libxenodon_b200.so")
CLI = os.path.join(PKG, "bin", "xenodon")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: the traversal kernels keep the reference shaders' binary32 operation order
# (no FMA contraction), see csrc/xn_device.cuh
NVCC_KERNEL_FLAGS = ["-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden"]
NVCC_HOST_FLAGS = ["-O2", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra", "-pthread"]

CU_SOURCES = [("xn_kernels.cu", NVCC_KERNEL_FLAGS), ("xn_util_kernels.cu", NVCC_HOST_FLAGS),
              ("xn_convert.cu", NVCC_HOST_FLAGS), ("xn_dag.cu", NVCC_HOST_FLAGS),
              ("xn_capi.cu", NVCC_HOST_FLAGS)]
CPP_SOURCES = ["host/xn_tiff.cpp", "host/xn_svo.cpp", "host/xn_text.cpp", "host/xn_png.cpp", "host/xn_synth_host.cpp"]
CLI_SOURCES = ["host/xn_cli.cpp"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _gxx() -> str:
    # the environment's CC/CXX point at a wrapper without OpenMP/pthread specs; use the system compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _sources_digest() -> str:
    h = hashlib.sha256()
    for base, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(base, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(ROOT, "include", "xenodon_b200.h"), "rb") as fh:
        h.update(fh.read())
    with open(os.path.abspath(__file__), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed:\n  " + " ".join(cmd) + "\n" + r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the shared library (and the CLI); returns the library path."""
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _sources_digest()
    if (not force and os.path.exists(LIB) and os.path.exists(CLI) and os.path.exists(stamp)
            and open(stamp).read() == digest):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    nvcc, gxx = _nvcc(), _gxx()
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    # tuning experiments: XN_NVCC_DEFS="-DXN_FAST_I2F=0" python -m xenodon_b200.build --force
    extra = os.environ.get("XN_NVCC_DEFS", "").split()
    objs = []
    for src, flags in CU_SOURCES:
        obj = os.path.join(OBJ, src.replace("/", "_") + ".o")
        cmd = [nvcc, "-ccbin", gxx, *ARCH, *flags, *extra, *inc, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
        objs.append(obj)
    for src in CPP_SOURCES:
        obj = os.path.join(OBJ, src.replace("/", "_") + ".o")
        cmd = [gxx, *CXX_FLAGS, *inc, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
        objs.append(obj)
    _run([nvcc, "-ccbin", gxx, *ARCH, "-shared", "-o", LIB, *objs, "-lz", "-Xlinker", "--no-undefined"])
    cli_objs = []
    for src in CLI_SOURCES:
        obj = os.path.join(OBJ, src.replace("/", "_") + ".o")
        _run([gxx, *CXX_FLAGS, *inc, "-c", os.path.join(CSRC, src), "-o", obj])
        cli_objs.append(obj)
    _run([gxx, "-o", CLI, *cli_objs, "-L", PKG, "-lxenodon_b200", "-Wl,-rpath,$ORIGIN/..", "-pthread"])
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


def build_variant(name: str, defs: str) -> str:
    """A/B experiments: libxenodon_b200 with xn_kernels.cu recompiled under extra -D flags, written to
    xenodon_b200/variants/libxenodon_b200_<name>.so (selected at run time with XN_LIBRARY=<path>)."""
    build()
    vdir = os.path.join(PKG, "variants")
    os.makedirs(vdir, exist_ok=True)
    nvcc, gxx = _nvcc(), _gxx()
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    obj = os.path.join(OBJ, f"xn_kernels_{name}.o")
    _run([nvcc, "-ccbin", gxx, *ARCH, *NVCC_KERNEL_FLAGS, *defs.split(), *inc, "-c", os.path.join(CSRC, "xn_kernels.cu"),
          "-o", obj])
    others = [os.path.join(OBJ, src.replace("/", "_") + ".o") for src, _ in CU_SOURCES if src != "xn_kernels.cu"]
    others += [os.path.join(OBJ, src.replace("/", "_") + ".o") for src in CPP_SOURCES]
    out = os.path.join(vdir, f"libxenodon_b200_{name}.so")
    _run([nvcc, "-ccbin", gxx, *ARCH, "-shared", "-o", out, obj, *others, "-lz", "-Xlinker", "--no-undefined"])
    os.remove(obj)  # 4 MB each, and build/ travels with every gpurun snapshot
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m xenodon_b200.build --variant NAME "-DXN_FOO=1 ..."
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2] if len(sys.argv) > i + 2 else ""))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose=True))
