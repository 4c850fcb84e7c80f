"""In-tree build of libhufb200.so (nvcc, sm_100a only).  No JIT cache: the .so sits next to
this file so that it travels with the repository snapshot to the GPU box."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# HUFB200_LIB: tuning sweeps load a prebuilt variant (never set in production)
LIB = os.environ.get("HUFB200_LIB") or os.path.join(HERE, "libhufb200.so")
SOURCES = ["huf_kernels.cu", "huf_api.cu"]
HEADERS = ["huf_device.cuh", "huf_kernels.h", os.path.join("..", "..", "include", "hufb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared", "-cudart", "shared",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libhufb200.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    # HUF_DEFS="-DHUF_DEC_BITS=10 ..." lets a tuning sweep rebuild variants (never set in production)
    extra = os.environ.get("HUF_DEFS", "").split()
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
