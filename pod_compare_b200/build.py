"""In-tree build of libpodb200.so (hand-written sm_100a CUDA behind a C ABI).

    python -m pod_compare_b200.build [--force] [-v]

nvcc cross-compiles for sm_100a without a GPU; the shared object lands next to this file so it
travels with the repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpodb200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["api.cu", "prep.cu", "conv_simt.cu", "conv_tc.cu", "score.cu", "decode.cu", "nms.cu", "merge.cu", "wire.cu", "backbone.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _extra_flags():
    """POD_LO_BITS=<0..10> in the environment rebuilds the library with that many mantissa bits in the lo operands of the
    fp16 split (csrc/common.cuh; default 7) -- for the accuracy / power experiments only."""
    v = os.environ.get("POD_LO_BITS")
    return ["-DPOD_LO_BITS=%d" % int(v)] if v not in (None, "") else []


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(_extra_flags()).encode())
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "podb200.h"))
    files.append(os.path.abspath(__file__))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    return "nvcc"


def _current(dig):
    if os.path.exists(OUT) and os.path.exists(STAMP):
        with open(STAMP) as f:
            return f.read().strip() == dig
    return False


def build(force=False, verbose=False):
    dig = _digest()
    if not force and _current(dig):
        return OUT
    # one builder at a time (torchrun starts several ranks at once)
    import fcntl
    lock = open(os.path.join(CSRC, ".build_lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and _current(dig):
            return OUT
        return _build_locked(dig, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(dig, verbose):
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc_path()] + NVCC_FLAGS + _extra_flags() + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s" % src)
    cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", OUT] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
