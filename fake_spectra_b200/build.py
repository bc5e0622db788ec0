"""Builds fake_spectra_b200/libfsb200.so (sm_100a only) in-tree with nvcc.

    python -m fake_spectra_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  cudart is linked statically so the library only needs the driver at run time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfsb200.so")
SOURCES = ["fsb_api.cu", "fsb_index.cu", "fsb_items.cu", "fsb_tau.cu", "fsb_colden.cu", "fsb_voronoi.cu", "fsb_stats.cu", "fsb_prep.cu", "fsb_microbench.cu"]
HEADERS = ["fsb_common.cuh", "fsb_scan.cuh", "fsb_voigt.cuh", "fsb_voigt_tables.h", "fsb_items.cuh"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def lib_path():
    return LIB


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "fsb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines, verbose=False):
    """Tuning aid: a second copy of the library with extra -D flags, selected at run time with
    FSB200_LIB=<path> (see _lib.py).  Returns the path of libfsb200_<name>.so."""
    return build(force=True, verbose=verbose, extra=["-D" + d for d in defines],
                 lib=os.path.join(HERE, "libfsb200_%s.so" % name), objsub="build_" + name)


def build(force=False, verbose=False, extra=(), lib=LIB, objsub="build"):
    if not force and not _stale():
        return LIB
    objdir = os.path.join(HERE, objsub)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([NVCC, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-o", lib] + objs, check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
