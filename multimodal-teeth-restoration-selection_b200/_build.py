"""Builds libteethrt.so (hand-written sm_100a CUDA behind the C ABI of include/teethrt.h) IN-TREE with nvcc.

nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container; the .so travels to the GPU box.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(PKG_DIR, "libteethrt.so")
SOURCES = ["abi.cu", "gemm_tc.cu", "eltwise.cu", "se_mlp.cu", "conv.cu", "dwconv.cu", "small.cu", "optim.cu", "preproc.cu", "calib.cu", "deskew.cu", "augment.cu"]
NVCC_FLAGS = (["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
              + os.environ.get("TEETHRT_NVCC_EXTRA", "").split())


# Pillow's float / double expressions must keep their operation order: no fused multiply-add contraction in this file
PER_FILE_FLAGS = {"augment.cu": ["--fmad=false"]}


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libteethrt cannot be built (there is no prebuilt or CPU fallback)")


def _deps_mtime():
    hdrs = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "augment_core.h"), os.path.join(PKG_DIR, "..", "include", "teethrt.h")]
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force):
    obj = os.path.join(BUILD, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), _deps_mtime()):
        return obj, False
    cmd = [_nvcc()] + NVCC_FLAGS + PER_FILE_FLAGS.get(src, []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, True


def build_library(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(lambda s: _compile(s, force), SOURCES))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(LIB_PATH):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print("built", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    build_library(force=True, verbose=True)
