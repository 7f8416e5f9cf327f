"""Build libadseis_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libadseis_b200%s.so" % os.environ.get("ADSEIS_LIB_SUFFIX", ""))  # suffix: tuning variants
SOURCES = ["ctx.cu", "acoustic.cu", "elastic.cu", "ops.cu"]
HEADERS = ["common.cuh", "acoustic_kernels.cuh", "elastic_kernels.cuh", "util_kernels.cuh", os.path.join("..", "..", "include", "adseis.h")]
# -fmad=false: fp64 products and sums are rounded separately, exactly like the reference's CPU op bodies, so that
# forward wavefields are bit-identical to the oracle.  The kernels are HBM-bound; the extra DMUL/DADD issue slots
# are not on the critical path.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS] + [__file__]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile if needed.  Safe under torchrun: ranks serialise on a file lock, the compiler writes to a temporary
    file and the result is moved into place atomically, so no rank can dlopen a half-written library.  A tuning
    variant (ADSEIS_LIB_SUFFIX set) that already exists is never rebuilt implicitly: its flags are not recorded."""
    if not force and os.environ.get("ADSEIS_LIB_SUFFIX") and os.path.exists(LIB):
        return LIB
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another rank built it while we waited
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            extra = os.environ.get("ADSEIS_NVCC_EXTRA", "").split()
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            env = dict(os.environ)
            env.pop("CC", None), env.pop("CXX", None)
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
            log = os.path.join(HERE, "build.log")
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed (see %s)" % log)
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
