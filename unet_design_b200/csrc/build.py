"""In-tree build of the sm_100a libraries (no JIT cache: the .so files travel with the repo snapshot).

  libunet_b200.so        C ABI (include/unet_b200.h), CUDA kernels only, static cudart, no torch
  libunet_b200_torch.so  TORCH_LIBRARY binding (torch.ops.unet_b200.*) over the C ABI

    python -m unet_design_b200.csrc.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
INCLUDE = os.path.join(ROOT, "include")
OUT_DIR = os.path.join(ROOT, "unet_design_b200")
LIB = os.path.join(OUT_DIR, "libunet_b200.so")
TORCH_LIB = os.path.join(OUT_DIR, "libunet_b200_torch.so")
OBJ_DIR = os.path.join(HERE, "build")

CU_SOURCES = ["api.cu", "haar.cu", "layout.cu", "groupnorm.cu", "optim.cu", "tc_host.cu", "conv_fprop.cu",
              "conv_wgrad.cu", "rowlin.cu", "attention.cu", "p2p.cu"]
HEADERS = ["common.cuh", "tc_common.cuh", os.path.join(INCLUDE, "unet_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", f"-I{INCLUDE}", f"-I{HERE}"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d if os.path.isabs(d) else os.path.join(HERE, d)) > t for d in deps)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_capi(force=False, verbose_ptxas=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs, procs = [], []
    for src in CU_SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [src] + HEADERS):
            cmd = [NVCC, *NVCC_FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
            if verbose_ptxas:
                cmd += ["-Xptxas", "-v"]
            print("+", " ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    if force or procs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-cudart", "static"])
    return LIB


def build_torch_binding(force=False):
    src = os.path.join(HERE, "torch_binding.cpp")
    if not (force or _newer(TORCH_LIB, [src, os.path.join(INCLUDE, "unet_b200.h")]) or
            os.path.getmtime(LIB) > os.path.getmtime(TORCH_LIB)):
        return TORCH_LIB
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{INCLUDE}", "-I/usr/local/cuda/include"]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, src, "-o", TORCH_LIB,
          f"-L{OUT_DIR}", "-lunet_b200", f"-L{torch_lib}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_cuda", "-lc10_cuda",
          "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"])
    return TORCH_LIB


def build_all(force=False):
    build_capi(force)
    build_torch_binding(force)


if __name__ == "__main__":
    build_capi(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
    if "--capi-only" not in sys.argv:
        build_torch_binding(force="--force" in sys.argv)
