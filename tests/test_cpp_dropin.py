"""The C++ drop-in header (include/mcmc_b200.hpp): compiles and links against the C-ABI library on the CPU; on a GPU
box the same program, written like reference user code, must reproduce the reference's golden draws bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import golden_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "bin")


def _build(src, out, std="c++14", defines=()):
    from mcmc_b200 import api

    assert os.path.exists(api.LIB_PATH), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    os.makedirs(BIN, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(api.LIB_PATH)
    cmd = [cxx, "-std=" + std, "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", out, "-L", libdir, "-lmcmc_b200",
           "-Wl,-rpath," + libdir] + ["-D" + x for x in defines]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def test_c_abi_header_is_plain_c_and_fails_loudly_without_a_gpu():
    """include/mcmc_b200.h + mcmc_b200_summary.h compile as C99 (what a cgo / JNI / ctypes binding sees); the example runs the
    hot path through the C ABI and, on a machine without a CUDA device, stops with the library's error — never a CPU result."""
    from mcmc_b200 import api

    os.makedirs(BIN, exist_ok=True)
    exe = os.path.join(BIN, "c_abi_hmc")
    libdir = os.path.dirname(api.LIB_PATH)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_hmc.c"),
                        "-o", exe, "-L", libdir, "-lmcmc_b200", "-Wl,-rpath," + libdir, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    if api.device_count() > 0:
        assert r.returncode == 0 and "acceptance rate" in r.stdout, (r.stdout, r.stderr)
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr, (r.stdout, r.stderr)


def test_draws_out_adapter_on_the_cpu():
    """b200_detail::unpack (chain-major device layout -> the reference's column-major Mat_t per chain, tiled + threaded)."""
    exe = _build(os.path.join(ROOT, "tests", "cpp", "unpack_check.cpp"), os.path.join(BIN, "unpack_check"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "unpack ok" in r.stdout, (r.stdout, r.stderr)


def test_header_and_examples_compile_and_link():
    _build(os.path.join(ROOT, "tests", "cpp", "dropin_check.cpp"), os.path.join(BIN, "dropin_check"), std="c++17")
    _build(os.path.join(ROOT, "examples", "hmc_normal.cpp"), os.path.join(BIN, "hmc_normal"))
    _build(os.path.join(ROOT, "examples", "rmhmc_funnel.cpp"), os.path.join(BIN, "rmhmc_funnel"))
    # the reference's fp32 build (MCMC_FPN_TYPE float): the same user source against fp_t compiles both ways, C++11 included
    _build(os.path.join(ROOT, "tests", "cpp", "fp32_check.cpp"), os.path.join(BIN, "fp32_check_f32"), std="c++11", defines=("MCMC_FPN_TYPE=float",))
    _build(os.path.join(ROOT, "tests", "cpp", "fp32_check.cpp"), os.path.join(BIN, "fp32_check_f64"), std="c++11")


def test_host_preconditioner_algebra_is_bit_identical_to_the_reference_operations():
    """tests/cpp/host_linalg_check.cpp: the inverse and the Cholesky factor the library computes on the host for precond_mat / cov_mat
    equal, bit for bit, A.inverse() and A.llt().matrixLLT() of the (stand-in) Eigen the reference is built against — 96 matrices, n up to 128."""
    from mcmc_b200 import api

    os.makedirs(BIN, exist_ok=True)
    exe = os.path.join(BIN, "host_linalg_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(api.LIB_PATH)
    r = subprocess.run([cxx, "-std=c++14", "-O2", "-ffp-contract=off", "-Wall", "-I", os.path.join(ROOT, "oracle", "standin"),
                        os.path.join(ROOT, "tests", "cpp", "host_linalg_check.cpp"), "-o", exe, "-L", libdir, "-lmcmc_b200", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "inverse_differs 0 cholesky_differs 0" in r.stdout, (r.stdout, r.stderr)


@pytest.mark.gpu
def test_dropin_program_reproduces_reference_goldens(engine):
    exe = _build(os.path.join(ROOT, "tests", "cpp", "dropin_check.cpp"), os.path.join(BIN, "dropin_check"), std="c++17")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    by = {c["name"]: c for c in golden_util.load()["cases"]}
    for c in golden_util.load()["de_cases"]:
        by[c["name"]] = dict(c, draws=np.array([float.fromhex(h) for h in c["draws_hex"]]).reshape(-1, c["draws_shape"][-1]))
    seen = set()
    for line in r.stdout.splitlines():
        tok = line.split()
        if tok[0] in by:
            rows, cols, acc = int(tok[1]), int(tok[2]), int(tok[3])
            draws = np.array([float.fromhex(h) for h in tok[4:]]).reshape(rows, cols)
            g = by[tok[0]]
            assert np.abs(draws - g["draws"]).max() <= 1e-10, tok[0]
            if tok[0] not in ("G4_rmhmc_normal", "hmc_box_d4"):  # device log()/exp() vs glibc in the Normal model / the transforms
                assert np.array_equal(draws, g["draws"]), tok[0]
            assert acc == g["n_accept"], tok[0]
            seen.add(tok[0])
        elif tok[0] == "multichain_consistent":
            assert tok[1] == "1" and tok[2] == "3"
            seen.add(tok[0])
        elif tok[0] == "sharded_consistent":
            assert tok[1] == "1"
            seen.add(tok[0])
        elif tok[0] in ("bounds_refused", "foreign_metric_refused"):
            assert tok[1] == "1"
            seen.add(tok[0])
    assert seen == {"G2_hmc_d3", "G3_mala_d3", "G4_rmhmc_normal", "rwmh_d3", "multichain_consistent", "sharded_consistent", "hmc_box_d4", "bounds_refused",
                    "G5_nuts_1d", "de_iso_d3", "foreign_metric_refused"}


@pytest.mark.gpu
def test_example_runs(engine):
    exe = _build(os.path.join(ROOT, "examples", "hmc_normal.cpp"), os.path.join(BIN, "hmc_normal"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    mean = [float(v) for v in r.stdout.splitlines()[0].split(":")[1].split("(")[0].split()]
    assert abs(mean[0] - 2.0) < 0.3 and abs(mean[1] - 2.0) < 0.3, r.stdout


@pytest.mark.gpu
def test_funnel_example_runs(engine):
    """BASELINE config 5 written like reference user code: RM-HMC, Neal's funnel d = 64, SoftAbs metric, many chains per call."""
    exe = _build(os.path.join(ROOT, "examples", "rmhmc_funnel.cpp"), os.path.join(BIN, "rmhmc_funnel"))
    r = subprocess.run([exe, "64"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    acc = float(r.stdout.strip().split("acceptance rate")[1])
    assert 0.2 < acc < 0.9, r.stdout


@pytest.mark.gpu
def test_fp32_build_of_the_header_returns_the_fp64_draws_narrowed(engine):
    """MCMC_FPN_TYPE float (include/misc/mcmc_options.hpp:80-99): ColVec_t / Mat_t / Cube_t, settings and target_data are fp32 at
    the boundary, the device computes in fp64.  One user source, built with and without -DMCMC_FPN_TYPE=float: every fp32
    draw is the fp64 draw rounded to float (the inputs are exactly representable), accept counts are identical."""
    src = os.path.join(ROOT, "tests", "cpp", "fp32_check.cpp")
    outs = {}
    for tag, defs in (("f32", ("MCMC_FPN_TYPE=float",)), ("f64", ())):
        exe = _build(src, os.path.join(BIN, "fp32_check_" + tag), std="c++11", defines=defs)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        d = {}
        for line in r.stdout.splitlines():
            tok = line.split()
            if tok[0] == "sizeof_fp_t":
                d["sizeof"] = int(tok[1])
            else:
                d[tok[0]] = (int(tok[3]), np.array([float.fromhex(h) for h in tok[4:]]).reshape(int(tok[1]), int(tok[2])))
        outs[tag] = d
    assert outs["f32"]["sizeof"] == 4 and outs["f64"]["sizeof"] == 8
    names = [k for k in outs["f64"] if k != "sizeof"]
    assert len(names) == 6
    for k in names:
        a32, d32 = outs["f32"][k]
        a64, d64 = outs["f64"][k]
        assert a32 == a64, k
        assert np.array_equal(d32, d64.astype(np.float32).astype(np.float64)), (k, np.abs(d32 - d64).max())
