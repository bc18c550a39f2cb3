"""In-tree build of the native code (sm_100a only).

  libdfa_b200.so                          CUDA kernels + the C ABI of include/dfa_b200.h (nvcc; four
                                          translation units under csrc/, compiled in parallel)
  ops/deformable_aggregation_ext*.so      thin torch extension with the reference's two entry
                                          points, forwarding raw pointers to the C ABI (g++)

Both land next to the sources so they travel with the repository snapshot to the GPU box.
"""
import os
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libdfa_b200.so")
KERNEL_SRCS = [os.path.join(PKG, "csrc", n) for n in ("dfa_forward.cu", "dfa_backward.cu", "dfa_frontend.cu",
                                                        "dfa_msda.cu")]
COMMON_HDR = os.path.join(PKG, "csrc", "dfa_common.cuh")
HEADERS = [os.path.join(PKG, "csrc", n) for n in ("dfa_common.cuh", "dfa_forward_win.cuh")]
OBJ_DIR = os.path.join(PKG, "csrc", "build")
EXT_SRC = os.path.join(PKG, "csrc", "dfa_torch_ext.cpp")
HEADER = os.path.join(ROOT, "include", "dfa_b200.h")
EXT = os.path.join(PKG, "ops", "deformable_aggregation_ext" + sysconfig.get_config_var("EXT_SUFFIX"))

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target, sources):
    return (not os.path.exists(target)
            or any(os.path.getmtime(target) < os.path.getmtime(s) for s in sources))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-4000:]))
    return r


def compile_objects(extra_flags=(), obj_dir=OBJ_DIR, verbose=False):
    """nvcc -c of every translation unit, in parallel; returns (object paths, ptxas stderr)."""
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in KERNEL_SRCS:
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = ["nvcc"] + NVCC_FLAGS + list(extra_flags) + ["-I", os.path.join(ROOT, "include"), "-c",
                                                           "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    objs, log = [], ""
    for obj, cmd, p in procs:
        out, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("build failed: %s\n%s%s" % (" ".join(cmd), out[-4000:], err[-4000:]))
        objs.append(obj)
        log += err
    return objs, log


def link_lib(objs, target=LIB):
    _run(["nvcc"] + LINK_FLAGS + ["-o", target] + objs)
    return target


def build_lib(force=False, verbose=False):
    if force or _stale(LIB, KERNEL_SRCS + HEADERS + [HEADER]):
        objs, log = compile_objects(verbose=verbose)
        link_lib(objs)
        if verbose:
            print(log, file=sys.stderr)
    return LIB


def build_ext(force=False):
    """The extension holds no device code: it validates tensors, takes PyTorch's current stream
    and calls libdfa_b200.so (found through an $ORIGIN rpath)."""
    build_lib(force)
    if force or _stale(EXT, [EXT_SRC, HEADER, LIB]):
        import torch
        from torch.utils import cpp_extension as ce
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", EXT_SRC, "-o", EXT,
               "-DTORCH_EXTENSION_NAME=deformable_aggregation_ext", "-DTORCH_API_INCLUDE_EXTENSION_H",
               "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
               "-I", os.path.join(ROOT, "include")]
        for p in ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]:
            cmd += ["-isystem", p]
        for d in ce.library_paths("cuda"):
            cmd += ["-L" + d, "-Wl,-rpath," + d]
        cmd += ["-L" + PKG, "-Wl,-rpath,$ORIGIN/..", "-ldfa_b200", "-lc10", "-lc10_cuda",
                "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
        _run(cmd)
    return EXT


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_ext(force)
    return LIB, EXT


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
