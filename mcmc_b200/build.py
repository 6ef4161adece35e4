"""Build mcmc_b200/libmcmc_b200.so (hand-written CUDA for sm_100a + the extern "C" shim).

Plain nvcc, no torch extension machinery: the product is a C-ABI shared library.
Each translation unit is compiled to an object in parallel, then linked with a
statically linked CUDA runtime so the .so runs next to (and independently of)
whatever libcudart a host process already has loaded.
"""
import hashlib
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmcmc_b200.so")

CU_SOURCES = ["engine.cu", "dispatch.cu", "hmc_wide.cu", "mala_wide.cu", "rmhmc.cu", "util_kernels.cu", "summary.cu", "rmhmc_general.cu",
              "hmc_batched.cu", "gather.cu", "transpose.cu", "rmhmc_cta.cu", "nuts_batched.cu", "hmc_duo.cu", "hmc_half.cu"]
# compiled once per registered target (-DMCMCB200_TARGET_SLICE=k), so the big template fan-out builds in parallel
SLICED_SOURCES = ["nuts.cu", "hmc.cu", "mala.cu", "rwmh.cu", "de.cu"]
# translation units that take longest (dense targets in the NUTS kernel: ~9 min each) start first
SLOW_FIRST = ["nuts.cu.t2.o", "nuts.cu.t3.o", "hmc.cu.t2.o", "hmc.cu.t3.o"]
N_TARGETS = 6
CPP_SOURCES = ["host_tape.cpp", "host_linalg.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
    "--compress-mode=size",   # the template fan-out is ~190 MB of SASS otherwise; the .so travels to every GPU box
    "-diag-suppress", "128",
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


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(path, seen=None):
    """The source and every project header it includes, transitively (so that touching one header only rebuilds its users)."""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in _INC.findall(f.read()):
            for base in (os.path.dirname(path), os.path.join(os.path.dirname(HERE), "include"), CSRC):
                if os.path.exists(os.path.join(base, inc)):
                    _deps(os.path.join(base, inc), seen)
                    break
    return seen


def build(verbose=False, force=False):
    """Set MCMCB200_FAST_BUILD=1 for a developer build with only the iso_gauss target (seconds instead of minutes);
    the default builds every registered target and is what __graft_entry__.build() and the tests use."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    fast = os.environ.get("MCMCB200_FAST_BUILD") == "1"
    units = []  # (source, object name, extra flags); the slowest translation units first
    for src in SLICED_SOURCES:
        for k in ([0] if fast else range(N_TARGETS)):
            units.append((src, "%s.t%d.o" % (src, k), ["-DMCMCB200_TARGET_SLICE=%d" % k]))
    units += [(src, src + ".o", []) for src in CU_SOURCES + CPP_SOURCES if os.path.exists(os.path.join(CSRC, src))]
    units.sort(key=lambda u: SLOW_FIRST.index(u[1]) if u[1] in SLOW_FIRST else len(SLOW_FIRST))
    base = NVCC_FLAGS + (["-DMCMCB200_FAST_BUILD"] if fast else [])
    jobs, objs = [], []
    for src, oname, extra in units:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, oname)
        objs.append(obj)
        flags = base + extra
        stamp_file = obj + ".stamp"
        stamp = _stamp(sorted(_deps(path)), flags)
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
               "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    build_user_example(verbose=verbose, force=force)
    return LIB


USER_EXAMPLE_SRC = os.path.join(os.path.dirname(HERE), "examples", "user_target", "normal_raw.cu")
USER_EXAMPLE_LIB = os.path.join(os.path.dirname(HERE), "examples", "user_target", "libnormal_raw.so")


def build_user_example(verbose=False, force=False):
    """examples/user_target/normal_raw.cu -> libnormal_raw.so: a USER-defined target built the way include/mcmc_b200_device.cuh
    documents (its own shared library next to libmcmc_b200.so; the library itself is not rebuilt)."""
    if not os.path.exists(USER_EXAMPLE_SRC):
        return None
    inc = os.path.join(os.path.dirname(HERE), "include")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "--compress-mode=size",
             "-diag-suppress", "128", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-I" + inc, "-I" + CSRC]
    stamp_file = USER_EXAMPLE_LIB + ".stamp"
    stamp = _stamp(sorted(_deps(USER_EXAMPLE_SRC) | _deps(os.path.join(inc, "mcmc_b200_register.cuh"))), flags)
    if not force and os.path.exists(USER_EXAMPLE_LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return USER_EXAMPLE_LIB
    cmd = [_nvcc(), "-ccbin", _host_cxx()] + flags + [USER_EXAMPLE_SRC, "-L" + HERE, "-lmcmc_b200", "-Xlinker", "-rpath=$ORIGIN/../../mcmc_b200",
                                                       "-o", USER_EXAMPLE_LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("user-target example failed to build:\n%s\n%s" % (r.stdout[-4000:], r.stderr[-8000:]))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return USER_EXAMPLE_LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
