"""Build recipe for the *unmodified* reference CUDA op (test infrastructure only).

Compiles the two native sources of the reference
(`projects/mmdet3d_plugin/ops/src/deformable_aggregation.cpp` and
`.../deformable_aggregation_cuda.cu`) from where they lie under /root/reference
into `oracle/_ref/deformable_aggregation_ext_ref*.so` for sm_100a.  Nothing is
copied into the repository: only the built shared object lands in `oracle/_ref/`
(git-ignored, but it travels to the GPU box with the snapshot).

The flags follow what `ops/setup.py:27-31` + torch's CUDAExtension would pass
(`-D__CUDA_NO_HALF_*`, default `-fmad=true`), with the arch pinned to sm_100a.

This module is imported only by `__graft_entry__.build()` (building the checker is
not using it), by `tests/` and by `bench.py`'s reference legs.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/projects/mmdet3d_plugin/ops/src"
# The reference binds its module as TORCH_EXTENSION_NAME; we give the built object a
# distinct name so it can be imported next to our own `deformable_aggregation_ext`.
MOD = "deformable_aggregation_ext_ref"


def so_path():
    return os.path.join(OUT, MOD + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=False):
    """Returns the path of the built .so, or None if the reference tree is absent
    (e.g. on the GPU box) and no prebuilt object exists."""
    target = so_path()
    srcs = [os.path.join(REF_SRC, "deformable_aggregation.cpp"),
            os.path.join(REF_SRC, "deformable_aggregation_cuda.cu")]
    if not all(os.path.exists(s) for s in srcs):
        return target if os.path.exists(target) else None
    if os.path.exists(target) and not force:
        if all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
            return target
    os.makedirs(OUT, exist_ok=True)
    import torch
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    common = ["-DTORCH_EXTENSION_NAME=" + MOD, "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    obj_cpp = os.path.join(OUT, "ref_binding.o")
    obj_cu = os.path.join(OUT, "ref_kernels.o")
    cmds = [
        ["g++", "-O3", "-std=c++17", "-fPIC", "-c", srcs[0], "-o", obj_cpp] + common + inc,
        ["nvcc", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-gencode", "arch=compute_100a,code=sm_100a",
         "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
         "-D__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr",
         "-c", srcs[1], "-o", obj_cu] + common + inc,
    ]
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", obj_cpp, obj_cu, "-o", target]
    for d in libdirs:
        link += ["-L" + d, "-Wl,-rpath," + d]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    cmds.append(link)
    import concurrent.futures as cf
    def run(c):
        r = subprocess.run(c, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed: %s\n%s" % (" ".join(c), r.stderr[-4000:]))
        if verbose:
            print(" ".join(c[:6]), "... ok", file=sys.stderr)
    with cf.ThreadPoolExecutor(2) as ex:
        list(ex.map(run, cmds[:2]))
    run(cmds[2])
    for o in (obj_cpp, obj_cu):
        os.remove(o)
    return target


def stage_python():
    """Copies the reference's Python files of the hot path (oracle/ref_import.py: FILES) from
    /root/reference into oracle/_ref/py — git-ignored like the built .so, and like it carried to the GPU
    box by the snapshot — so that bench.py's reference arm can time the reference's OWN
    feature_sampling + multi_view_level_fusion (models/blocks.py:215-261) there.  Returns the staged
    root, or None when neither the reference tree nor an earlier copy exists."""
    from oracle import ref_import
    src_root = "/root/reference"
    if ref_import.available(src_root):
        for f in ref_import.FILES:
            dst = os.path.join(ref_import.STAGED_ROOT, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(src_root, f), dst)
    return ref_import.STAGED_ROOT if ref_import.available(ref_import.STAGED_ROOT) else None


def load():
    """Import the built reference extension (needs a CUDA device to *run*)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    p = so_path()
    if not os.path.exists(p):
        raise FileNotFoundError(p + " — run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(MOD, p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    print(build(force="--force" in sys.argv, verbose=True))
    print(stage_python())
