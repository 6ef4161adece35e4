"""Build mcmc_b200/libmcmc_b200.so (hand-written CUDA for sm_100a + the extern "C" shim).

Plain nvcc, no torch extension machinery: the product is a C-ABI shared library.
Each translation unit is compiled to an object in parallel, then linked with a
statically linked CUDA runtime so the .so runs next to (and independently of)
whatever libcudart a host process already has loaded.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmcmc_b200.so")

CU_SOURCES = ["engine.cu", "dispatch.cu", "hmc_wide.cu", "mala_wide.cu", "rmhmc.cu", "util_kernels.cu", "summary.cu", "rmhmc_general.cu"]
# compiled once per registered target (-DMCMCB200_TARGET_SLICE=k), so the big template fan-out builds in parallel
SLICED_SOURCES = ["hmc.cu", "nuts.cu", "mala.cu", "rwmh.cu"]
N_TARGETS = 6
CPP_SOURCES = ["host_tape.cpp", "host_linalg.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stamp(paths, extra):
    h = hashlib.sha256()
    h.update(repr(extra).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "mcmc_b200.h"))
    return hs


def build(verbose=False, force=False):
    """Set MCMCB200_FAST_BUILD=1 for a developer build with only the iso_gauss target (seconds instead of minutes);
    the default builds every registered target and is what __graft_entry__.build() and the tests use."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = _headers()
    fast = os.environ.get("MCMCB200_FAST_BUILD") == "1"
    units = []  # (source, object name, extra flags); the slowest translation units first
    for src in SLICED_SOURCES:
        for k in ([0] if fast else range(N_TARGETS)):
            units.append((src, "%s.t%d.o" % (src, k), ["-DMCMCB200_TARGET_SLICE=%d" % k]))
    units += [(src, src + ".o", []) for src in CU_SOURCES + CPP_SOURCES]
    base = NVCC_FLAGS + (["-DMCMCB200_FAST_BUILD"] if fast else [])
    jobs, objs = [], []
    for src, oname, extra in units:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, oname)
        objs.append(obj)
        flags = base + extra
        stamp_file = obj + ".stamp"
        stamp = _stamp([path] + headers, flags)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        cmd = [nvcc, "-ccbin", _host_cxx()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        jobs.append((cmd, stamp_file, stamp, oname))

    def run(job):
        cmd, stamp_file, stamp, src = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout[-4000:], r.stderr[-8000:]))
        with open(stamp_file, "w") as f:
            f.write(stamp)
        return src, r.stderr

    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 4, max(1, len(jobs)))) as ex:
        for src, log in ex.map(run, jobs):
            if verbose:
                sys.stderr.write("== %s\n%s\n" % (src, log))

    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-ccbin", _host_cxx(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
