#!/usr/bin/env python
"""Build the reference's own (legacy Deformable-DETR) CUDA extension for sm_100a as a GPU baseline.

Runs in the authoring container only (needs /root/reference).  The reference sources are copied to
a scratch directory under /tmp, two tokens are patched so they compile against torch 2.11
(`value.type()` -> `value.scalar_type()` at ms_deform_attn_cuda.cu:69,139), and the
extension is compiled with torch.utils.cpp_extension.  Only the resulting shared object is kept,
under the git-ignored baseline/_ref/ (it travels to the GPU box with the snapshot); no reference
source enters the repository.  Nothing under pctrans_b200/ imports it: it is loaded by
scripts/bench_legacy.py alone, as "the kernel to beat".
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OPS_SRC = "/root/reference/connectomics/model/maskformer_block/pixel_decoder/ops/src"
OUT_DIR = os.path.join(HERE, "_ref")
SCRATCH = "/tmp/pct_legacy_msda"


def main():
    if not os.path.isdir(OPS_SRC):
        print("reference not present; keeping any prebuilt baseline/_ref as is")
        return 0
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    shutil.rmtree(SCRATCH, ignore_errors=True)
    shutil.copytree(OPS_SRC, os.path.join(SCRATCH, "src"))
    cu = os.path.join(SCRATCH, "src", "cuda", "ms_deform_attn_cuda.cu")
    text = open(cu).read()
    assert text.count("AT_DISPATCH_FLOATING_TYPES(value.type(),") == 2
    open(cu, "w").write(text.replace("AT_DISPATCH_FLOATING_TYPES(value.type(),",
                                     "AT_DISPATCH_FLOATING_TYPES(value.scalar_type(),"))
    from torch.utils.cpp_extension import load

    src = os.path.join(SCRATCH, "src")
    sources = [os.path.join(src, "vision.cpp")] + glob.glob(os.path.join(src, "cpu", "*.cpp")) + \
        glob.glob(os.path.join(src, "cuda", "*.cu"))
    build_dir = os.path.join(SCRATCH, "build")
    os.makedirs(build_dir, exist_ok=True)
    load(name="MultiScaleDeformableAttention", sources=sources, extra_include_paths=[src],
         extra_cflags=["-DWITH_CUDA"],
         extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                            "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                            "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=build_dir, with_cuda=True, is_python_module=False, verbose=False)
    os.makedirs(OUT_DIR, exist_ok=True)
    so = os.path.join(build_dir, "MultiScaleDeformableAttention.so")
    shutil.copy(so, os.path.join(OUT_DIR, "MultiScaleDeformableAttention.so"))
    print("built", os.path.join(OUT_DIR, "MultiScaleDeformableAttention.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
