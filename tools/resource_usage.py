"""Registers / stack / shared memory of the hot kernels, read with cuobjdump from the objects build() leaves in
mcmc_b200/build (CPU-only).  Output committed as profiles/r2_resource_usage.txt; tests/test_cabi_cpu.py guards the headline kernel."""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ["hmc_pipe_kernel", "hmc_half_kernel", "hmc_duo_kernel", "hmc_wide", "mala_rows_kernel", "dgemm_dmma", "nuts_pc_kernel", "nuts_ls_step",
       "nuts_ls_gemm", "nuts_coop", "rmhmc_cta_kernel", "hmc_batched_rows", "chain_stats_kernel"]


def usage(obj):
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
    rows, fn = [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            rows.append((fn, int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
            fn = None
    return rows


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return p.stdout.splitlines()


def main():
    rows = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "mcmc_b200", "build", "*.o"))):
        for r in usage(obj):
            if any(h in r[0] for h in HOT):
                rows.append((os.path.basename(obj),) + r)
    names = demangle([r[1] for r in rows])
    print("%-22s %4s %6s %7s %6s  %s" % ("object", "regs", "stack", "shared", "local", "kernel"))
    for r, n in zip(rows, names):
        n = re.sub(r"\(anonymous namespace\)::", "", n).replace("mcmcb200::", "")
        n = re.sub(r"\(.*\)$", "", n).replace("void ", "")
        print("%-22s %4d %6d %7d %6d  %s" % (r[0], r[2], r[3], r[4], r[5], n))


if __name__ == "__main__":
    main()
