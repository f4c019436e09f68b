"""Build `libjolideco_b200.so` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m jolideco_b200.build [--force] [--verbose]

Every `csrc/*.cu` is compiled to its own object (in parallel, cached by a digest of the source, the headers and
the flags under `jolideco_b200/build/`), then linked into the shared library.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
# experiment builds (tools/): JD_NVCC_EXTRA="-DJD_X=1" JD_LIB_TAG=x python -m jolideco_b200.build writes
# libjolideco_b200_x.so next to the product library; JD_LIB_PATH selects it at load time (_lib.py)
TAG = os.environ.get("JD_LIB_TAG", "")
LIB = os.path.join(HERE, f"libjolideco_b200{'_' + TAG if TAG else ''}.so")
STAMP = os.path.join(HERE, f".build_stamp{'_' + TAG if TAG else ''}")
OBJDIR = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
] + os.environ.get("JD_NVCC_EXTRA", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(INCLUDE, "jolideco_b200.h"))
    return files


def _hash(files, extra=""):
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(extra.encode())
    return h.hexdigest()


def _digest():
    return _hash(sources() + _headers())


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, f"{os.path.basename(src)[:-3]}.{_hash([src] + _headers())[:16]}.o")
    if os.path.exists(obj):
        return obj, ""
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj + ".tmp"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{res.stdout}{res.stderr}")
    os.replace(obj + ".tmp", obj)
    return obj, res.stdout + res.stderr


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as pool:
        results = list(pool.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    nvcc = os.environ.get("NVCC", "nvcc")
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libjolideco_b200.so")
    keep = set(objs)
    for f in os.listdir(OBJDIR):  # drop stale objects
        if os.path.join(OBJDIR, f) not in keep:
            os.remove(os.path.join(OBJDIR, f))
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
