"""CPU-only probe: the stand-in Eigen's dense product (what the CPU baseline runs) beside OpenBLAS dgemv on the same core.
Builds tools/standin_gemv_probe.cpp with the baseline's flags (oracle/Makefile FAST, single thread), times the C2 hot-loop
expressions at d = 128, then times scipy's dgemv (net of the Python call overhead, measured with a 4 x 4 product).
The ratio bounds how much faster a real-Eigen build of the reference could run C2's dense algebra.  Measurement tooling."""
import json
import os
import subprocess
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    exe = os.path.join(ROOT, "tools", "bin", "standin_gemv_probe")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++14", "-O3", "-march=x86-64-v3", "-ffp-contract=fast", "-DNDEBUG", "-I" + os.path.join(ROOT, "oracle", "standin"),
                    os.path.join(ROOT, "tools", "standin_gemv_probe.cpp"), "-o", exe], check=True)
    standin = json.loads(subprocess.run([exe, "128", "200000"], check=True, capture_output=True, text=True).stdout)
    from scipy.linalg.blas import dgemv
    from threadpoolctl import threadpool_limits

    per_call = {}
    with threadpool_limits(1):
        for d in (4, 128):
            M, p, y = np.asfortranarray(np.eye(d)), np.arange(1, d + 1) * 1e-3, np.zeros(d)
            for _ in range(2000):
                dgemv(1.0, M, p, y=y, overwrite_y=1)
            n = 300000
            t = time.perf_counter()
            for _ in range(n):
                dgemv(1.0, M, p, y=y, overwrite_y=1)
            per_call[d] = (time.perf_counter() - t) / n
    net = per_call[128] - per_call[4]
    blas = 2 * 128 * 128 / net * 1e-9
    cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]
    print(json.dumps(dict(cpu=cpu, standin=standin, openblas_dgemv_128=dict(us_net=net * 1e6, gflops=blas),
                          bound_on_real_eigen_speedup_of_the_dense_part=blas / standin["gflops"])))


if __name__ == "__main__":
    main()
