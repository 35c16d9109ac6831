"""Build the UNMODIFIED reference extensions into oracle/_ref/ (test infrastructure only).

The five torch-ngp extensions that Seal-3D ships (raymarching, gridencoder, shencoder,
freqencoder, ffmlp) are compiled from the sources *where they lie* under /root/reference
(nothing is copied into this repo) for sm_100a, under private module names
``_ref_raymarching`` ... so they can never shadow the product modules.  The only deviation
from the reference's own flags (``*/backend.py``) is ``-std=c++17`` (torch >= 2.1 headers
reject c++14) and the explicit ``-gencode arch=compute_100a,code=sm_100a``.

The resulting ``oracle/_ref/_ref_*.so`` files are git-ignored but travel to the GPU box with
gpurun; there they are the *primary* witness for parity (reference kernels, same GPU) and
are used to (re)generate ``tests/golden/*.npz``.  They are never imported by the product
package -- only by tests/ and by tests/golden/make_gpu_golden.py.

Usage:  python oracle/build_ref.py [name ...]      (run in the build container; needs
        /root/reference; several minutes per extension)
"""
import os
import sys
import subprocess

REF = os.environ.get("SEAL3D_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

NVCC_COMMON = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
]

EXTS = {
    "raymarching": dict(srcs=["raymarching.cu", "bindings.cpp"], nvcc=[], inc=[]),
    "gridencoder": dict(srcs=["gridencoder.cu", "bindings.cpp"], nvcc=[], inc=[]),
    "shencoder": dict(srcs=["shencoder.cu", "bindings.cpp"], nvcc=[], inc=[]),
    "freqencoder": dict(srcs=["freqencoder.cu", "bindings.cpp"], nvcc=["-use_fast_math"], inc=[]),
    "ffmlp": dict(
        srcs=["ffmlp.cu", "bindings.cpp"],
        nvcc=["--expt-extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler=-mf16c",
              "-Xcompiler=-Wno-float-conversion", "-Xcompiler=-fno-strict-aliasing"],
        inc=["dependencies/cutlass/include", "dependencies/cutlass/tools/util/include"],
    ),
}


def build_one(name):
    from torch.utils.cpp_extension import load
    cfg = EXTS[name]
    src_dir = os.path.join(REF, name, "src")
    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    # freqencoder's reference backend.py uses -use_fast_math; check and mirror
    load(
        name="_ref_" + name,
        sources=[os.path.join(src_dir, s) for s in cfg["srcs"]],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=NVCC_COMMON + cfg["nvcc"],
        extra_include_paths=[os.path.join(REF, name, p) for p in cfg["inc"]],
        build_directory=bdir,
        is_python_module=False,
        verbose=True,
    )
    so = os.path.join(bdir, "_ref_%s.so" % name)
    dst = os.path.join(OUT, "_ref_%s.so" % name)
    if os.path.exists(so):
        import shutil
        shutil.copy2(so, dst)
    print("built", dst)


WRAPPERS = {"raymarching": ["__init__.py", "raymarching.py"], "gridencoder": ["__init__.py", "grid.py"],
            "shencoder": ["__init__.py", "sphere_harmonics.py"], "freqencoder": ["__init__.py", "freq.py"], "ffmlp": ["__init__.py", "ffmlp.py"]}


def stage_wrappers():
    """The reference's Python wrappers (gridencoder/grid.py, raymarching/raymarching.py, ...) are the other half of the
    extension surface: `nerf/network.py` calls THEM, and they call `_gridencoder` / `_raymarching`.  Like the compiled
    extensions they cannot be read from /root/reference on the GPU box, so the unmodified files are staged next to the
    built `.so` files under the git-ignored oracle/_ref/py/ (never committed, never imported by the product); the GPU test
    tests/test_gpu_reference_wrappers.py runs them once over the reference's extensions and once over this repo's."""
    import shutil
    for pkg, files in WRAPPERS.items():
        dst = os.path.join(OUT, "py", pkg)
        os.makedirs(dst, exist_ok=True)
        for f in files:
            shutil.copy2(os.path.join(REF, pkg, f), os.path.join(dst, f))
    print("staged the reference wrappers under", os.path.join(OUT, "py"))


def main():
    names = sys.argv[1:] or list(EXTS)
    if not os.path.isdir(REF):
        print("reference tree %s absent: nothing to build (prebuilt oracle/_ref/*.so are used as-is)" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    if names == ["wrappers"]:
        stage_wrappers()
        return 0
    if len(sys.argv) == 1:
        stage_wrappers()
    if len(names) == 1:
        build_one(names[0])
        return 0
    procs = [(n, subprocess.Popen([sys.executable, os.path.abspath(__file__), n],
                                  stdout=open(os.path.join(OUT, "build_%s.log" % n), "w"),
                                  stderr=subprocess.STDOUT)) for n in names]
    rc = 0
    for n, p in procs:
        r = p.wait()
        print(n, "rc=%d" % r)
        rc |= r
    return rc


if __name__ == "__main__":
    sys.exit(main())
