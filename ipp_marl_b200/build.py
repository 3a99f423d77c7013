"""Build the CUDA library (nvcc, sm_100a) in-tree: ipp_marl_b200/_lib/libipp_b200.so.

The built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libipp_b200.so")
STAMP = os.path.join(LIBDIR, "libipp_b200.stamp")
SOURCES = ["ipp_kernels.cu", "ipp_step_tma.cu", "ipp_features.cu", "ipp_planner.cu", "ipp_facade_kernels.cu", "ipp_abi.cu"]
# every header a source may include: all of csrc/*.cuh, csrc/*.h and the public ABI header
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
    os.path.join("..", "..", "include", "ipp_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
] + os.environ.get("IPP_NVCC_EXTRA", "").split()  # development only (timing knobs: scripts/dbg_bench.py)


def _digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    return src, res.returncode, res.stdout, res.stderr


def _up_to_date(digest):
    if os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            return f.read().strip() == digest
    return False


def build(force=False, verbose=False):
    """Compile if sources changed; returns the library path.  Raises if nvcc fails.

    Each .cu is compiled to its own object (in parallel, cached by content digest under _lib/obj/) and the
    objects are linked into one shared library.  Concurrent callers (one process per GPU under torchrun) are
    serialised by a file lock: the first one builds, the others find the library up to date."""
    os.makedirs(LIBDIR, exist_ok=True)
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB
    import fcntl

    with open(os.path.join(LIBDIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and _up_to_date(digest):  # another process built it while we waited
            return LIB
        return _build_locked(force, verbose, digest)


def _build_locked(force, verbose, digest):
    nvcc = nvcc_path()
    if nvcc is None:
        if os.path.exists(LIB):
            # box without a toolkit: the prebuilt library shipped with the snapshot must match these sources
            raise RuntimeError("prebuilt %s does not match the sources (stamp mismatch) and nvcc is not available "
                               "to rebuild it" % LIB)
        raise RuntimeError("nvcc not found and no prebuilt %s" % LIB)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    hh = hashlib.sha256()
    for name in HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            hh.update(f.read())
    hh.update(" ".join(NVCC_FLAGS).encode())
    jobs, objs = [], []
    for src in SOURCES:
        h = hh.copy()
        with open(os.path.join(CSRC, src), "rb") as f:
            h.update(f.read())
        obj = os.path.join(objdir, "%s.%s.o" % (src[:-3], h.hexdigest()[:16]))
        objs.append(obj)
        if force or verbose or not os.path.exists(obj):
            for stale in os.listdir(objdir):
                if stale.startswith(src[:-3] + "."):
                    os.remove(os.path.join(objdir, stale))
            jobs.append((nvcc, src, obj, verbose))
    if jobs:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            for src, rc, out, err in pool.map(_compile_one, jobs):
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (src, out, err))
                if verbose:
                    print(err)
    res = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
