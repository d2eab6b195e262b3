"""In-tree build of the C-ABI CUDA library (nvcc cross-compiles for sm_100a without a GPU).

    python -m hypernet_image_captioning_b200.build [--force]

Output: hypernet_image_captioning_b200/lib/libcaphn_b200.so (git-ignored, travels to the GPU box with the snapshot).
"""
import glob
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libcaphn_b200.so")
STAMP = LIB + ".srchash"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def source_hash():
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == source_hash()


def _file_hash(path, common):
    h = hashlib.sha256(common)
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu (one nvcc process per file, in parallel; objects cached per source hash under lib/obj/)
    and link them into one shared library; no-op when the sources are unchanged."""
    if not force and is_current():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    ch = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sorted(glob.glob(os.path.join(CSRC, "*.cuh"))):
        with open(f, "rb") as fh:
            ch.update(fh.read())
    common = ch.digest()
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".hash"
        hv = _file_hash(src, common)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == hv:
            return obj, None
        cmd = [_nvcc()] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            return obj, "nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr
        if verbose:
            sys.stderr.write(res.stderr)
        with open(stamp, "w") as fh:
            fh.write(hv)
        return obj, None

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, srcs))
    errs = [e for _, e in results if e]
    if errs:
        raise RuntimeError("\n".join(errs))
    # No -lcuda: the one driver-API symbol needed (cuTensorMapEncodeTiled) is fetched with cudaGetDriverEntryPoint.
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + \
        [o for o, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
