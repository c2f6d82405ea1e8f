"""In-tree build of libhyperpose_b200.so (nvcc, sm_100a only).  `python -m hyperpose_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhyperpose_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall", "-Xptxas", "-v"]
# bit-exact fp32 kernels (parser): never let nvcc contract a*b+c
EXACT_FLAGS = ["-fmad=false"]

# (source, extra flags)
SOURCES = [
    ("paf_parser.cu", EXACT_FLAGS),
    ("engine.cu", []),
    ("pifpaf_decoder.cu", EXACT_FLAGS),
    ("ppn_parser.cu", EXACT_FLAGS),
    ("common.cpp", []),
    ("handoff.cpp", []),
    ("pool.cpp", []),
]


def _newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "hyperpose_b200.h"))
    objs = []
    rebuilt = False
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _newer([s] + headers, o):
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
            with open(o + ".ptxas.txt", "w") as f:
                f.write(r.stderr)
            rebuilt = True
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs  # static cudart (nvcc default); the driver API is reached via cudaGetDriverEntryPoint
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


def build_cpp_example(ref_root: str = "/root/reference") -> str | None:
    """Compiles the C++ drop-in (hyperpose_api/*.cpp) + examples/operator_api_b200.cpp against the reference's
    UNCHANGED public headers.  Needs the reference tree (headers are not copied into this repo); returns the
    binary path, or None when the reference is absent (GPU box: the prebuilt binary travels with the snapshot)."""
    exe = os.path.join(ROOT, "examples", "operator_api_b200")
    if not os.path.isdir(os.path.join(ref_root, "include", "hyperpose")):
        return exe if os.path.exists(exe) else None
    api = os.path.join(CSRC, "hyperpose_api")
    srcs = [os.path.join(api, "paf.cpp"), os.path.join(api, "tensorrt.cpp"), os.path.join(api, "pifpaf.cpp"), os.path.join(api, "pose_proposal.cpp"),
            os.path.join(ROOT, "examples", "operator_api_b200.cpp")]
    if _newer(srcs + [LIB], exe):
        cmd = ["g++", "-std=c++17", "-O2", "-DHP_B200_STANDALONE", "-I" + os.path.join(CSRC, "shim"), "-I" + os.path.join(ref_root, "include"),
               "-I" + os.path.join(ROOT, "include")] + srcs + ["-L" + PKG, "-lhyperpose_b200", "-Wl,-rpath,$ORIGIN/../hyperpose_b200", "-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("C++ drop-in failed to compile against the reference headers")
    return exe


def build_stream_example(ref_root: str = "/root/reference", mock: bool = False) -> str | None:
    """examples/stream_api_b200.cpp: the reference's STREAM API on the drop-in.  The scheduler is the reference's own --
    include/hyperpose/stream/stream.hpp instantiated as it is, src/stream.cpp + src/thread_pool.cpp + src/logging.cpp compiled from
    the reference tree unchanged.  mock=True builds the CPU self-check variant (stand-in engine / parser, no GPU, no library)."""
    exe = os.path.join(ROOT, "examples", "stream_api_b200_mock" if mock else "stream_api_b200")
    if not os.path.isdir(os.path.join(ref_root, "include", "hyperpose")):
        return exe if os.path.exists(exe) else None
    api = os.path.join(CSRC, "hyperpose_api")
    ref = [os.path.join(ref_root, "src", f) for f in ("stream.cpp", "thread_pool.cpp", "logging.cpp")]
    srcs = [os.path.join(ROOT, "examples", "stream_api_b200.cpp")] + ref
    if not mock:
        srcs += [os.path.join(api, "paf.cpp"), os.path.join(api, "tensorrt.cpp")]
    if _newer([s for s in srcs if s.startswith(ROOT)] + ([] if mock else [LIB]) + [os.path.join(CSRC, "shim", "opencv2", "opencv.hpp")], exe):
        cmd = ["g++", "-std=c++17", "-O2", "-pthread", "-DHP_STREAM_MOCK" if mock else "-DHP_B200_STANDALONE", "-I" + os.path.join(CSRC, "shim"),
               "-I" + os.path.join(ref_root, "include"), "-I" + os.path.join(ref_root, "src"), "-I" + os.path.join(ROOT, "include")] + srcs
        if not mock:
            cmd += ["-L" + PKG, "-lhyperpose_b200", "-Wl,-rpath,$ORIGIN/../hyperpose_b200"]
        cmd += ["-o", exe]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("stream example failed to compile against the reference's stream sources")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
