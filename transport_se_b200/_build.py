"""In-tree builds: host mesh library (g++) and the CUDA C-ABI library (nvcc, sm_100a)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
ROOT = os.path.dirname(_HERE)


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_host(force=False):
    out = os.path.join(_HERE, "libtse_host.so")
    srcs = [os.path.join(CSRC, "tse_mesh.cpp"), os.path.join(CSRC, "tse_mesh.hpp")]
    if force or _newer(out, srcs):
        _run(["g++", "-O2", "-std=gnu++17", "-fopenmp", "-fPIC", "-shared", "-o", out, srcs[0], "-lquadmath"])
    return out


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force=False, verbose=False, out=None):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -> libtse_cuda.so (cross-compiles without a GPU)."""
    out = out or os.path.join(_HERE, "libtse_cuda.so")
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))]
    deps.append(os.path.join(ROOT, "include", "tse.h"))
    if force or _newer(out, deps):
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += os.environ.get("TSE_NVCC_FLAGS", "").split()
        cmd += ["-o", out] + srcs + ["-lnccl"]
        log = _run(cmd)
        if verbose:
            print(log)
    return out


def build_oracle(force=False):
    """Test infrastructure only (see oracle/oracle.cpp header)."""
    odir = os.path.join(ROOT, "oracle")
    out = os.path.join(odir, "_build", "liboracle.so")
    src = os.path.join(odir, "oracle.cpp")
    if force or _newer(out, [src]):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        _run(["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src])
    return out


def build_driver(force=False):
    """driver/prim_main: the C++ stand-alone driver (namelist, time loop, printstate, error norms) linked against libtse_cuda.so."""
    out = os.path.join(ROOT, "driver", "prim_main")
    srcs = [os.path.join(ROOT, "driver", "prim_main.cpp"), os.path.join(CSRC, "tse_mesh.cpp")]
    deps = srcs + [os.path.join(CSRC, "tse_mesh.hpp"), os.path.join(ROOT, "include", "tse.h"), os.path.join(_HERE, "libtse_cuda.so")]
    if force or _newer(out, deps):
        _run(["g++", "-O2", "-std=gnu++17", "-fopenmp", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", out] + srcs +
             ["-L", _HERE, "-ltse_cuda", "-lquadmath", "-Wl,-rpath," + _HERE, "-Wl,--allow-shlib-undefined"])
    return out


def build_c_abi_check(force=False):
    """tests/c_abi/abi_check: a C11 program that includes include/tse.h and drives the library without ctypes."""
    out = os.path.join(ROOT, "tests", "c_abi", "abi_check")
    src = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")
    deps = [src, os.path.join(ROOT, "include", "tse.h"), os.path.join(_HERE, "libtse_cuda.so"), os.path.join(_HERE, "libtse_host.so")]
    if force or _newer(out, deps):
        _run(["gcc", "-O1", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", out, src,
              "-L", _HERE, "-ltse_cuda", "-ltse_host", "-lm", "-Wl,-rpath," + _HERE, "-Wl,--allow-shlib-undefined"])
    return out
