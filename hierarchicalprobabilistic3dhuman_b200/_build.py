"""In-tree build of libhp3d.so (nvcc, sm_100a only). Used by __graft_entry__.build()."""
import os
import subprocess
import shutil

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libhp3d.so")
HOST_SHIM_PATH = os.path.join(PKG_DIR, "libhp3d_hostshim.so")
SOURCES = ["api.cu", "smpl.cu", "smpl_fused.cu", "mf_sampler.cu", "mf_head.cu", "encoder.cu", "conv_tc.cu", "gemm_tc.cu", "rank.cu", "proxy.cu", "crop.cu", "mf_norm.cu", "peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(PKG_DIR, "..", "include", "hp3d.h"))
    return d


def build(force=False, verbose=False):
    """Compile every CUDA source into libhp3d.so (cross-compiles without a GPU) and the host-side
    SVD shim used by the CPU tests. Returns the library path."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if force or not _newer(LIB_PATH, _deps()):
        # one nvcc process per translation unit, in parallel (no relocatable device code: every kernel is self-contained),
        # then a link step; objects live in build/ (git-ignored)
        from concurrent.futures import ThreadPoolExecutor
        obj_dir = os.path.join(PKG_DIR, "build")
        os.makedirs(obj_dir, exist_ok=True)
        flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

        def compile_one(src):
            obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
            r = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)
            return obj, r

        with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as ex:
            results = list(ex.map(compile_one, srcs))
        for obj, r in results:
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            if verbose:
                print(r.stderr)
        r = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH]
                           + [obj for obj, _ in results], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    shim_src = os.path.join(CSRC, "host_shim.cpp")
    if force or not _newer(HOST_SHIM_PATH, [shim_src, os.path.join(CSRC, "svd3.h"), os.path.join(CSRC, "crop_math.h"), os.path.join(CSRC, "mf_norm_math.h")]):
        r = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", HOST_SHIM_PATH, shim_src],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return LIB_PATH
