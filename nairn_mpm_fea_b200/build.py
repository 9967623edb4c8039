"""Build libmpmgpu.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and the tests."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmpmgpu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared"]


def nccl_include():
    """nccl.h for the TYPES of the run-time-resolved NCCL calls in capi.cu (the system's, else the wheel torch depends on)."""
    for d in ("/usr/include", os.path.join(os.path.dirname(os.path.dirname(shutil.which("python") or "")), "lib")):
        if os.path.exists(os.path.join(d, "nccl.h")):
            return ["-I" + d]
    try:
        import nvidia.nccl
        return ["-I" + os.path.join(list(nvidia.nccl.__path__)[0], "include")]
    except Exception:
        return []


def sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            out.append(os.path.join(CSRC, f))
    out.append(os.path.join(os.path.dirname(HERE), "include", "mpmgpu.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """Compile csrc/capi.cu (which includes every kernel header) into libmpmgpu.so."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("MPMGPU_NVCC_DEFS", "").split()        # e.g. "-DF2_MINB=5" for tuning runs
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, "capi.cu"), "-o", LIB, "-ldl"] + nccl_include()
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (p.stdout, p.stderr))
    if verbose:
        print(p.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
