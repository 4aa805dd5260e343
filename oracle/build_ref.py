"""Compile the reference's own native code as a witness for the oracle -- TEST INFRASTRUCTURE ONLY.

Sources are compiled where they lie under /root/reference (never copied into the repo); outputs go
only to oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the built .so files travel to the GPU box).
On the GPU box /root/reference does not exist: the prebuilt modules are imported as they are.

  dss_ref_cpu  : DSS/csrc/rasterize_points_cpu.cpp through oracle/ref_shim_cpu.cpp          (CPU)
  dss_ref_frnn_cpu : external/FRNN/frnn/csrc/bruteforce/bruteforce_cpu.cpp (K-NN ground truth)   (CPU)
  dss_ref_cuda : DSS/csrc/rasterize_points.cu, rasterize_points_backward.cu,
                 external/prefix_sum/prefix_sum.cu, external/FRNN/frnn/csrc/grid/{grid,counting_sort}.cu
                 through oracle/ref_shim_cuda.cpp, nvcc sm_100a                              (GPU witness)
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference"


def _load_prebuilt(name):
    path = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (the module links against libtorch)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stale(name, sources):
    so = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(so):
        return True
    t = os.path.getmtime(so)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build_cpu(verbose=False):
    srcs = [os.path.join(HERE, "ref_shim_cpu.cpp"), os.path.join(REF, "DSS/csrc/rasterize_points_cpu.cpp")]
    if not os.path.isdir(REF) or not _stale("dss_ref_cpu", srcs):
        return _load_prebuilt("dss_ref_cpu")
    from torch.utils.cpp_extension import load
    bd = os.path.join(OUT, "dss_ref_cpu")
    os.makedirs(bd, exist_ok=True)
    return load(name="dss_ref_cpu", sources=srcs, build_directory=bd, extra_cflags=["-O2"],
                extra_include_paths=[os.path.join(REF, "DSS/csrc")], verbose=verbose)


def build_cuda(verbose=False):
    srcs = [os.path.join(HERE, "ref_shim_cuda.cpp"),
            os.path.join(REF, "DSS/csrc/rasterize_points.cu"),
            os.path.join(REF, "DSS/csrc/rasterize_points_backward.cu"),
            os.path.join(REF, "external/FRNN/frnn/csrc/grid/grid.cu"),
            os.path.join(REF, "external/FRNN/frnn/csrc/grid/counting_sort.cu")]
    if not os.path.isdir(REF) or not _stale("dss_ref_cuda", srcs):
        return _load_prebuilt("dss_ref_cuda")
    from torch.utils.cpp_extension import load
    bd = os.path.join(OUT, "dss_ref_cuda")
    os.makedirs(bd, exist_ok=True)
    stub = os.path.join(OUT, "stub")       # empty THC/THCNumerics.cuh (rasterize_points.cu:5), SURVEY App. C
    os.makedirs(os.path.join(stub, "THC"), exist_ok=True)
    for f in ("THCNumerics.cuh",):
        p = os.path.join(stub, "THC", f)
        if not os.path.exists(p):
            open(p, "w").write("// stub: header removed from recent PyTorch; nothing from it is used\n")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    return load(name="dss_ref_cuda", sources=srcs, build_directory=bd,
                extra_cflags=["-O2", "-DWITH_CUDA"],
                extra_cuda_cflags=["-O3", "-DWITH_CUDA", "-gencode", "arch=compute_100a,code=sm_100a",
                                   "--expt-relaxed-constexpr", "-lineinfo"],
                extra_include_paths=[os.path.join(REF, "DSS/csrc"), stub,
                                     os.path.join(REF, "external/prefix_sum"),
                                     os.path.join(REF, "external/FRNN/frnn/csrc")],
                verbose=verbose)


def build_prefix_sum(verbose=False):
    """external/prefix_sum/prefix_sum.cu as its own module: its header carries its own PYBIND11_MODULE
    (external/prefix_sum/prefix_sum.h:18-21) exposing prefix_sum_cuda / prefix_sum_cpu."""
    srcs = [os.path.join(REF, "external/prefix_sum/prefix_sum.cu")]
    if not os.path.isdir(REF) or not _stale("dss_ref_prefix_sum", srcs):
        return _load_prebuilt("dss_ref_prefix_sum")
    from torch.utils.cpp_extension import load
    bd = os.path.join(OUT, "dss_ref_prefix_sum")
    os.makedirs(bd, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    return load(name="dss_ref_prefix_sum", sources=srcs, build_directory=bd,
                extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"], verbose=verbose)


def build_frnn_cpu(verbose=False):
    """external/FRNN/frnn/csrc/bruteforce/bruteforce_cpu.cpp through oracle/ref_shim_frnn_cpu.cpp."""
    srcs = [os.path.join(HERE, "ref_shim_frnn_cpu.cpp"),
            os.path.join(REF, "external/FRNN/frnn/csrc/bruteforce/bruteforce_cpu.cpp")]
    if not os.path.isdir(REF) or not _stale("dss_ref_frnn_cpu", srcs):
        return _load_prebuilt("dss_ref_frnn_cpu")
    from torch.utils.cpp_extension import load
    bd = os.path.join(OUT, "dss_ref_frnn_cpu")
    os.makedirs(bd, exist_ok=True)
    return load(name="dss_ref_frnn_cpu", sources=srcs, build_directory=bd, extra_cflags=["-O2"], verbose=verbose)


def ref_frnn_cpu():
    try:
        return build_frnn_cpu()
    except Exception as e:  # pragma: no cover
        print("oracle/_ref frnn cpu build failed:", e, file=sys.stderr)
        return None


def ref_cpu():
    """The reference CPU module, or None when neither /root/reference nor a prebuilt .so exists."""
    try:
        return build_cpu()
    except Exception as e:  # pragma: no cover
        print("oracle/_ref cpu build failed:", e, file=sys.stderr)
        return None


def ref_cuda():
    try:
        return build_cuda()
    except Exception as e:  # pragma: no cover
        print("oracle/_ref cuda build failed:", e, file=sys.stderr)
        return None


def ref_prefix_sum():
    try:
        return build_prefix_sum()
    except Exception as e:  # pragma: no cover
        print("oracle/_ref prefix_sum build failed:", e, file=sys.stderr)
        return None


def build_all(verbose=False):
    return build_cpu(verbose), build_cuda(verbose), build_prefix_sum(verbose), build_frnn_cpu(verbose)


if __name__ == "__main__":
    print(build_all(verbose=True))
