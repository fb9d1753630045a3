"""Build recipe for libdcc_b200.so (nvcc, sm_100a only, in-tree so the .so travels with the repo)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(PKG, "libdcc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


HASH_PATH = LIB_PATH + ".srchash"


def source_hash():
    """Content hash of everything the library is compiled from (sources, headers, flags).  Staleness is decided on
    content, not mtimes: a snapshot copied to another box (gpurun) keeps its prebuilt .so whatever the copy did to the
    timestamps, and an edited source always triggers a rebuild."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libdcc_b200.so cannot be built (no fallback path exists)")
    return nvcc


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:      # one builder at a time (torchrun starts N ranks at once)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():               # another rank built it while we waited
            return LIB_PATH
        return _build_locked(verbose)


def _build_locked(verbose):
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd += ["-o", tmp] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
