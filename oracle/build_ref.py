"""Oracle A: compile the reference's OWN CUDA rasterizer and simple-knn for sm_100a.

TEST INFRASTRUCTURE. Nothing under oracle/ is imported by the product path; only tests/, bench.py
(`--impl reference` and the cpu_baseline leg) and __graft_entry__.smoke() may use it, as the checker.

The reference sources are compiled WHERE THEY LIE under /root/reference (never copied into this
repo); the only outputs are two torch-extension shared objects under oracle/_ref/ (git-ignored,
but shipped to the GPU box by gpurun):

    oracle/_ref/ref_dgr_C.so   <- submodules_local/diff-gaussian-rasterization/{cuda_rasterizer/*.cu,
                                   rasterize_points.cu, ext.cpp}   (pybind module, ext.cpp:15-19)
    oracle/_ref/ref_knn_C.so   <- submodules_local/simple-knn/{simple_knn.cu, spatial.cu, ext.cpp}

Accommodations: GLM is an un-vendored submodule of the reference, so
oracle/glm_standin/glm/glm.hpp is put on the include path; modern libstdc++ needs `-include cstdint`
(rasterizer_impl.h) and `-include cfloat` (simple_knn.cu:90). Nothing else is changed.

This script is a no-op (returns False) when /root/reference is absent, i.e. on the GPU box, which
only uses the prebuilt files.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GSR_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _torch_flags():
    from torch.utils.cpp_extension import include_paths, library_paths

    inc = ["-I" + p for p in include_paths()]
    inc.append("-I" + sysconfig.get_paths()["include"])
    libs = ["-L" + p for p in library_paths()]
    libs += ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda"]
    rpath = []
    for p in library_paths():
        rpath += ["-Xlinker", "-rpath", "-Xlinker", p]
    import torch

    abi = ["-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    return inc, libs + rpath, abi


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _build(name, srcs, extra, verbose):
    """One `nvcc -c` per source, run in parallel (torch/CUB headers make each take minutes), then link."""
    from concurrent.futures import ThreadPoolExecutor

    inc, link, abi = _torch_flags()
    target = os.path.join(OUT, name + ".so")
    deps = srcs + [os.path.join(HERE, "glm_standin", "glm", "glm.hpp"), os.path.abspath(__file__)]
    if _newer(target, deps):
        return target
    objdir = os.path.join(HERE, "_build", name)
    os.makedirs(objdir, exist_ok=True)
    common = (
        ["-std=c++17", "-O3", "-Xcompiler", "-fPIC", "-lineinfo"]
        + ARCH
        + ["--expt-relaxed-constexpr", "-w", "-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H"]
        + abi
        + extra
        + inc
    )

    def one(i_src):
        i, src = i_src
        obj = os.path.join(objdir, "%d_%s.o" % (i, os.path.splitext(os.path.basename(src))[0]))
        cmd = ["nvcc"] + common + ["-x", "cu", "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=max(1, min(len(srcs), os.cpu_count() or 1))) as ex:
        objs = list(ex.map(one, enumerate(srcs)))
    cmd = ["nvcc", "-shared"] + ARCH + objs + ["-o", target] + link
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return target


# The reference's Python CALLERS of the path (gaussian_renderer.render(), the model and camera classes, the losses) and its own
# autograd wrapper, copied verbatim into the git-ignored baseline/_ref/refpy/ (shipped to the GPU box by
# gpurun, never committed). tests/test_reference_callers_gpu.py imports them UNCHANGED -- once over this repo's drop-in
# `diff_gaussian_rasterization`, once over the reference's own wrapper + ref_dgr_C.so -- and compares everything they return.
CALLER_FILES = [
    "gaussian_renderer/__init__.py", "scene/cameras.py", "scene/gaussian_model.py", "utils/graphics_utils.py", "utils/general_utils.py",
    "utils/sh_utils.py", "utils/system_utils.py", "utils/loss_utils.py",
    "submodules_local/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py",
]
CALLERS_OUT = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "refpy")


def copy_callers():
    import shutil

    if not os.path.isdir(REF):
        return False
    for rel in CALLER_FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(CALLERS_OUT, rel.replace("submodules_local/diff-gaussian-rasterization/", "ref_wrapper/"))
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
            shutil.copyfile(src, dst)
    return True


def build(verbose=False):
    if not os.path.isdir(REF):
        return False
    copy_callers()
    os.makedirs(OUT, exist_ok=True)
    dgr = os.path.join(REF, "submodules_local", "diff-gaussian-rasterization")
    knn = os.path.join(REF, "submodules_local", "simple-knn")
    from concurrent.futures import ThreadPoolExecutor

    jobs = [
        ("ref_dgr_C",
         [os.path.join(dgr, "cuda_rasterizer", "rasterizer_impl.cu"), os.path.join(dgr, "cuda_rasterizer", "forward.cu"),
          os.path.join(dgr, "cuda_rasterizer", "backward.cu"), os.path.join(dgr, "rasterize_points.cu"), os.path.join(dgr, "ext.cpp")],
         ["-include", "cstdint", "-I" + os.path.join(HERE, "glm_standin"), "-I" + dgr]),
        ("ref_knn_C",
         [os.path.join(knn, "simple_knn.cu"), os.path.join(knn, "spatial.cu"), os.path.join(knn, "ext.cpp")],
         ["-include", "cfloat", "-include", "cstdint", "-I" + knn]),
    ]
    with ThreadPoolExecutor(max_workers=2) as ex:  # both modules at once; each compiles its sources in parallel too
        list(ex.map(lambda j: _build(j[0], j[1], j[2], verbose), jobs))
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref built" if ok else "reference tree not present: nothing built")
    sys.exit(0)
