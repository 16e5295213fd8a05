"""Build the reference's own CUDA ops, unmodified, for sm_100a -> oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing here is imported by the product path.

The reference sources are compiled *where they lie* under /root/reference (nothing is
copied into this repo); only the build products land in oracle/_ref/ (git-ignored, but
shipped to the GPU box by gpurun).  Two extension modules result:

  oracle/_ref/chamfer/chamfer.so        <- extensions/chamfer_dist/{chamfer.cu,chamfer_cuda.cpp}
  oracle/_ref/pointnet2_ext/_ext.so     <- extensions/pointnet2/_ext_src/src/*.{cpp,cu}

  oracle/_ref/pysrc/*.py                <- byte copies of the reference's own Python that sits directly on those
                                           ops (extensions/pointnet2/pointnet2_utils.py), staged so the GPU box -- which has
                                           no /root/reference -- can run the reference module over the real _ext.so and
                                           produce GPU-made golden vectors (tests/golden/make_golden_pointnet2.py --gpu).
                                           oracle/_ref/ is git-ignored: no reference source enters the history.

They are the *live GPU oracle* (run on the gpurun box by tests/ and by
tests/golden/make_golden.py) and the "reference CUDA recompiled for B200" speed comparator
in bench.py's `ref_gpu` block.  They cannot run in the build container (no GPU).

Usage:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("POINTDAE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def _load(name, sources, include_dirs, build_dir, verbose):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils import cpp_extension

    os.makedirs(build_dir, exist_ok=True)
    # is_python_module=True imports the product once, which proves it links.
    return cpp_extension.load(
        name=name,
        sources=sources,
        extra_include_paths=include_dirs,
        extra_cuda_cflags=["-O2", "-gencode", "arch=compute_100a,code=sm_100a"],
        extra_cflags=["-O2"],
        build_directory=build_dir,
        verbose=verbose,
    )


def build(verbose=False):
    """Returns dict name -> path of the built .so (empty when the reference is absent)."""
    built = {}
    if not os.path.isdir(REF):
        return built
    ch = os.path.join(REF, "extensions", "chamfer_dist")
    p2 = os.path.join(REF, "extensions", "pointnet2", "_ext_src")
    jobs = [
        ("chamfer", [os.path.join(ch, "chamfer_cuda.cpp"), os.path.join(ch, "chamfer.cu")], [],
         os.path.join(OUT, "chamfer")),
        ("_ext", sorted(glob.glob(os.path.join(p2, "src", "*.cpp")) + glob.glob(os.path.join(p2, "src", "*.cu"))),
         [os.path.join(p2, "include")], os.path.join(OUT, "pointnet2_ext")),
    ]
    for name, srcs, incs, bdir in jobs:
        so = os.path.join(bdir, name + ".so")
        if not os.path.exists(so):
            _load(name, srcs, incs, bdir, verbose)
        built[name] = so
    built.update(stage_python())
    return built


PY_SOURCES = {"pointnet2_utils.py": ("extensions", "pointnet2", "pointnet2_utils.py")}


def stage_python():
    """Copies (never edits) the reference Python listed in PY_SOURCES into oracle/_ref/pysrc/ (git-ignored)."""
    import shutil
    out = {}
    dst_dir = os.path.join(OUT, "pysrc")
    os.makedirs(dst_dir, exist_ok=True)
    for name, parts in PY_SOURCES.items():
        src = os.path.join(REF, *parts)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(dst_dir, name))
            out["pysrc/" + name] = os.path.join(dst_dir, name)
    return out


if __name__ == "__main__":
    out = build(verbose="-v" in sys.argv)
    for k, v in out.items():
        print(k, "->", v, "(exists)" if os.path.exists(v) else "(MISSING)")
