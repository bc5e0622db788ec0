"""TEST INFRASTRUCTURE — builds the two CPU checkers used by tests/, smoke() and bench.py.

1. ``oracle/_build/libfsoracle.so``  — the C restatement (oracle/fs_oracle.c), always buildable.
2. ``oracle/_ref/libfsref.so``       — the UNMODIFIED reference hot path, compiled from the sources
   where they lie under /root/reference/fake_spectra (absorption.cpp, index_table.cpp,
   part_int.cpp, Faddeeva.cpp) plus oracle/ref_shim.cpp.  Flags follow the reference's
   Makefile:10 (``-O3 -ffast-math -fopenmp``).  Only built when /root/reference exists (i.e. in
   the build container); the GPU box uses the prebuilt .so that travels with the snapshot.
3. ``oracle/_ref/ref_boost_test`` / ``ref_faddeeva_test`` — the reference's own test.cpp (against
   a minimal Boost.Test shim, oracle/boost_shim) and Faddeeva self-test, run by
   ``python oracle/build.py --selftest``.

No reference source is copied into the repo; outputs go to oracle/_ref/ (git-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("FSB_REFERENCE_SRC", "/root/reference/fake_spectra")
REF_OUT = os.path.join(HERE, "_ref")
ORACLE_OUT = os.path.join(HERE, "_build")
REF_FILES = ["absorption.cpp", "index_table.cpp", "part_int.cpp", "Faddeeva.cpp"]
REF_FLAGS = ["-O3", "-g", "-fPIC", "-ffast-math", "-fopenmp"]  # reference Makefile:10


def _run(cmd):
    subprocess.run(cmd, check=True)


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.exists(s) and os.path.getmtime(s) <= t for s in sources)


def oracle_lib_path():
    return os.path.join(ORACLE_OUT, "libfsoracle.so")


def ref_lib_path():
    return os.path.join(REF_OUT, "libfsref.so")


def have_reference_sources():
    return all(os.path.exists(os.path.join(REF_SRC, f)) for f in REF_FILES)


def build_oracle(force=False):
    """Compile the C restatement.  Strict IEEE (no -ffast-math) so that it is a fixed point."""
    src = os.path.join(HERE, "fs_oracle.c")
    out = oracle_lib_path()
    if not force and _newer(out, [src]):
        return out
    os.makedirs(ORACLE_OUT, exist_ok=True)
    _run(["gcc", "-O2", "-std=c11", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", src, "-o", out, "-lm"])
    return out


def build_ref(force=False):
    """Compile the unmodified reference sources + shim into oracle/_ref/libfsref.so."""
    out = ref_lib_path()
    if not have_reference_sources():
        return out if os.path.exists(out) else None
    srcs = [os.path.join(REF_SRC, f) for f in REF_FILES] + [os.path.join(HERE, "ref_shim.cpp")]
    if not force and _newer(out, srcs):
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    _run(["g++"] + REF_FLAGS + ["-shared", "-I", REF_SRC] + srcs + ["-o", out])
    return out


def build_ref_selftests():
    """Reference test.cpp (unmodified, Boost shim) and the Faddeeva self-test."""
    if not have_reference_sources():
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    btest = os.path.join(REF_OUT, "ref_boost_test")
    _run(["g++", "-O3", "-ffast-math", "-fopenmp", "-I", os.path.join(HERE, "boost_shim"), "-I", REF_SRC,
          os.path.join(REF_SRC, "test.cpp"), os.path.join(REF_SRC, "absorption.cpp"),
          os.path.join(REF_SRC, "index_table.cpp"), os.path.join(REF_SRC, "Faddeeva.cpp"), "-o", btest])
    ftest = os.path.join(REF_OUT, "ref_faddeeva_test")
    _run(["g++", "-O2", "-DTEST_FADDEEVA", os.path.join(REF_SRC, "Faddeeva.cpp"), "-o", ftest])
    return btest, ftest


def build_all(force=False):
    return build_oracle(force), build_ref(force)


if __name__ == "__main__":
    o, r = build_all(force="--force" in sys.argv)
    print("oracle:", o)
    print("reference:", r)
    if "--selftest" in sys.argv:
        tests = build_ref_selftests()
        if tests:
            for t in tests:
                res = subprocess.run([t], capture_output=True, text=True)
                tail = (res.stdout + res.stderr).strip().splitlines()[-3:]
                print(os.path.basename(t), "exit", res.returncode, "|", " | ".join(tail))
