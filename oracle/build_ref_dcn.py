"""Compile the reference's OWN CUDA extension `deform_conv_ext` (code/models/modules/ops/dcn/src/*.cpp|.cu, the pybind11 module that
deform_conv.py:23-26 imports) for sm_100, from the sources where they lie under /root/reference, into oracle/_ref/ (git-ignored, travels to
the GPU box with the repo snapshot).  TEST / MEASUREMENT INFRASTRUCTURE: it is the "reference GPU kernel as shipped" that tools/gpu/dcn_ref_compare.py
times beside glare's DCN kernels and that tests/test_dcn_gpu.py uses as a second, reference-built checker.  No reference source is copied.

    python -m oracle.build_ref_dcn          # authoring container only (needs /root/reference); ~3 min
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/code/models/modules/ops/dcn/src"
OUT = os.path.join(HERE, "_ref")


def build(verbose=False):
    if not os.path.isdir(SRC):
        raise RuntimeError("reference tree not mounted (%s)" % SRC)
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    return load("deform_conv_ext", [os.path.join(SRC, f) for f in ("deform_conv_ext.cpp", "deform_conv_cuda.cpp", "deform_conv_cuda_kernel.cu")],
                build_directory=OUT, verbose=verbose, is_python_module=False)


if __name__ == "__main__":
    build(verbose="-v" in sys.argv)
    print(sorted(os.listdir(OUT)))
