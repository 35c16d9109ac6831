"""In-tree build of the CUDA library and the reference-compatible extension modules.

  libseal3d_b200.so   csrc/*.cu  -> the C-ABI of include/seal3d_b200.h (no torch, no Python)
  _raymarching.so, _gridencoder.so, _shencoder.so, _freqencoder.so, _ffmlp.so
                      csrc/shims/*.cpp -> pybind11/torch modules with the reference's module names
                      and signatures (raymarching/src/bindings.cpp etc.), thin wrappers that unpack
                      at::Tensor and call the C-ABI on the current torch stream.

Everything is compiled for sm_100a only (``-gencode arch=compute_100a,code=sm_100a -lineinfo``);
nvcc cross-compiles without a GPU.  Outputs stay in-tree (git-ignored) so they travel with gpurun.
"""
import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libseal3d_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "-I", CSRC, "-I", INCLUDE] + os.environ.get("S3D_NVCC_EXTRA", "").split()   # e.g. -DS3D_TRACE (scripts/trace_bwd.py)
SHIMS = ["raymarching", "gridencoder", "shencoder", "freqencoder", "ffmlp"]


def _host_cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stamp(paths, extra=""):
    h = hashlib.sha1(extra.encode())
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_lib(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = cuda_sources()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)] if os.path.isdir(INCLUDE) else []
    hstamp = _stamp(hdrs, " ".join(NVCC_FLAGS))

    def one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        st = _stamp([src], hstamp)
        stf = obj + ".stamp"
        if not force and os.path.exists(obj) and os.path.exists(stf) and open(stf).read() == st:
            return obj, False
        _run([NVCC, "-ccbin", _host_cxx()] + NVCC_FLAGS + ["-c", src, "-o", obj], verbose)
        open(stf, "w").write(st)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(one, srcs))
    objs = [o for o, _ in res]
    if force or any(ch for _, ch in res) or not os.path.exists(LIB):
        _run([NVCC, "-ccbin", _host_cxx(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs +
             ["-lcudart"], verbose)
    return LIB


def build_shims(verbose=False, force=False):
    """The five pybind modules.  Needs torch headers; ~1 min of g++ per module the first time."""
    import torch
    from torch.utils import cpp_extension as ce
    shim_dir = os.path.join(CSRC, "shims")
    inc = ce.include_paths()
    pyinc = sysconfig.get_paths()["include"]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = "/usr/local/cuda/include"
    os.makedirs(OBJ, exist_ok=True)

    def one(name):
        src = os.path.join(shim_dir, name + ".cpp")
        out = os.path.join(HERE, "_%s.so" % name)
        st = _stamp([src, os.path.join(shim_dir, "shim_common.h"), os.path.join(INCLUDE, "seal3d_b200.h")], torch.__version__)
        stf = os.path.join(OBJ, "_%s.stamp" % name)
        if not force and os.path.exists(out) and os.path.exists(stf) and open(stf).read() == st:
            return out
        cmd = [_host_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_%s" % name,
               "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
               "-I", INCLUDE, "-I", pyinc, "-I", cuda_inc]
        for p in inc:
            cmd += ["-isystem", p]
        cmd += [src, "-o", out, "-L", HERE, "-lseal3d_b200", "-Wl,-rpath,$ORIGIN", "-L", tlib, "-ltorch", "-ltorch_cpu",
                "-ltorch_python", "-lc10", "-lc10_cuda", "-ltorch_cuda", "-Wl,-rpath," + tlib,
                "-L/usr/local/cuda/lib64", "-lcudart"]
        _run(cmd, verbose)
        open(stf, "w").write(st)
        return out

    with ThreadPoolExecutor(max_workers=5) as ex:
        return list(ex.map(one, SHIMS))


def build_all(verbose=False, force=False, shims=True):
    lib = build_lib(verbose, force)
    outs = [lib]
    if shims:
        outs += build_shims(verbose, force)
    return outs


if __name__ == "__main__":
    v = "-q" not in sys.argv
    print(build_all(verbose=v, force="--force" in sys.argv, shims="--no-shims" not in sys.argv))
