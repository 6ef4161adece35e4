#!/usr/bin/env python
"""bench.py — the headline measurement of BASELINE.json: HMC draws/sec (chains x iterations, d=128).

A "step" is one pass of the hot path over one batch of synthetic input: the whole C2 job of SURVEY.md §8(d) —
mcmc::hmc semantics on log pi(x) = -|x|^2/2, d=128, 4096 chains PER GPU, L=10, eps=0.1, M=I, fp64,
n_burnin=100 + n_keep=1000 draws, x0[c][j]=sin(0.37c+0.11j), in-kernel Philox (seed 12345).  Chains shard
across ranks with no data-path collective (weak scaling: 4096 chains per GPU, global chain ids).

  value    draws/s of the whole job with inputs resident in HBM (device pointers through the C ABI)
  e2e      the same metric through the C ABI with HOST (pinned) buffers: H2D of x0 and D2H of draws_out and
           n_accept_draws inside the timed region; d2h_ceiling = a bare cudaMemcpy of the same bytes at the same time on
           every rank (what the host's PCIe / memory system allows), e2e.frac_of_ceiling relates the two
  e2e_cpp  the same job through the reference-shaped C++ call mcmc::hmc(Mat_t initial_vals, kernel, Cube_t& draws_out, ...)
           of include/mcmc_b200.hpp (tools/e2e_cpp.cpp, built by __graft_entry__.build()), rank 0 only
  roofline algorithmic bytes (2*d*8 B per transition) / CUDA-event kernel time vs the measured HBM peak
  strong   BASELINE north_star's target line: the SAME 4096 chains split over the N ranks (512 per GPU at N=8)
  value_with_gather / allgather: weak-scaling steps that also assemble draws_out on every rank (the library's NCCL
           all-gather, mcmcb200_allgather_draws) inside the timed region
  configs  the other BASELINE configs, chains sharded over the ranks: c3 (MALA d=1024 linreg, 16384 chains), c4 (NUTS
           d=256 dense Gaussian cond 1e3, 4096 chains), c5 (RM-HMC funnel d=64 SoftAbs, 2048 chains), sweep (HMC
           iso-Gaussian d in {32,128,512,2048}, 4096 chains) — each with kernel time, its roofline and (N=1) a CPU leg
  cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref, OpenMP loop over chains, one mcmc::hmc
           call per chain) on the host cores, bounded sample of the same workload.
  e2e_summary (informational, not the reference's output format): the same call chain when the caller only needs
           posterior summaries — draws_out stays in HBM and mcmcb200_summarize_draws reduces it there.
With --gpus N > 1 every rank first binds itself to the host cores of its GPU's NUMA node (best effort).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, CHAINS_PER_GPU, N_BURNIN, N_KEEP, LEAP, EPS, SEED = 128, 4096, 100, 1000, 10, 0.1, 12345
N_TOTAL = N_BURNIN + N_KEEP
ALG_BYTES_PER_DRAW = 2 * D * 8  # read x_prev + write x_new (= draws_out row), SURVEY.md §8(d)
WORKLOAD = ("C2: mcmc::hmc, iso-Gaussian d=128, %d chains/GPU, L=%d, eps=%g, M=I, fp64, %d burn-in + %d kept draws"
            % (CHAINS_PER_GPU, LEAP, EPS, N_BURNIN, N_KEEP))
L2_FLUSH_BYTES = 256 << 20  # 2x the 126 MB L2


def initial_vals(first_chain, n_chains, d=D):
    import numpy as np

    c = np.arange(first_chain, first_chain + n_chains, dtype=np.float64)[:, None]
    j = np.arange(d, dtype=np.float64)[None, :]
    return np.sin(0.37 * c + 0.11 * j)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            busy = sorted(s for s, pw in zip(sm, power) if pw > 0.5 * max(power)) or sorted(sm)
            out.update(sm_mhz=busy[len(busy) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def ref_lib():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol

    if ol.Reference.available("fast"):
        return ol, ol.Reference("fast"), "reference"
    return ol, None, "port"


def bind_to_gpu_numa_node(local_rank, world):
    """Run this rank on the host cores nearest to its GPU (so the pinned staging buffers of the end-to-end leg are
    allocated there and the D2H stream does not cross the socket interconnect).  The GPU's NUMA node comes from sysfs; when
    the platform does not report it (numa_node = -1, as on the round-1 scaling box) the allowed CPUs are split into `world`
    contiguous slices and the rank takes its own — ranks then at least do not compete for the same cores.  Best effort: a
    missing piece leaves the affinity untouched.  Returns a short description."""
    try:
        import torch

        allowed_all = sorted(os.sched_getaffinity(0))
        node = -1
        try:
            p = torch.cuda.get_device_properties(local_rank)
            bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        except (OSError, ValueError):
            node = -1
        if node >= 0:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = set(allowed_all) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                return "numa node %d (%d cpus)" % (node, len(allowed))
        per = len(allowed_all) // max(1, world)
        if per >= 1:
            mine = allowed_all[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, set(mine))
            return "numa node not reported: cpu slice %d-%d of the allowed set (%d cpus)" % (mine[0], mine[-1], len(mine))
        return "numa node not reported; too few cpus to slice"
    except Exception as e:  # noqa: BLE001
        return "not bound (%s)" % type(e).__name__


def usable_cpus():
    """Host threads this process may really use: affinity mask, capped by the cgroup CPU quota if there is one."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = min(n, max(1, int(float(q[0]) / float(q[1]))))
    except (OSError, ValueError, IndexError):
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return max(1, n)


def cpu_reference_throughput(steps=1, warmup=0, target_seconds=8.0):
    """draws/s of the unmodified reference (oracle/_ref: OpenMP loop over chains, one mcmc::hmc call per chain) on all
    usable host threads.  One 'step' = a bounded sample of the C2 workload: n_chains chains x N_TOTAL draws with the
    same d, L, eps and seeds, n_chains sized from a short pilot so that a step takes about target_seconds."""
    ol, ref, kind = ref_lib()
    cores = usable_cpus()

    def run(n_chains, n_burnin, n_keep):
        st = ol.Settings(n_burnin=n_burnin, n_keep=n_keep, n_leap_steps=LEAP, step_size=EPS)
        x0 = initial_vals(0, n_chains)
        if ref is not None:
            return ref.run_chains(ol.HMC, ol.TGT_ISO_GAUSS, None, x0, st, SEED, n_threads=cores, keep_draws=False)[2]
        orc = ol.Oracle()  # reference could not be compiled here: time the restated port, single thread
        t0 = time.perf_counter()
        for c in range(n_chains):
            orc.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, x0[c], st, seed=SEED + c)
        return time.perf_counter() - t0

    if ref is None:
        cores = 1
    pilot_chains = max(2, cores)
    run(pilot_chains, 10, 100)                            # first call also starts the OpenMP thread pool
    el = run(pilot_chains, N_BURNIN, N_KEEP)              # pilot: one full-length chain per thread (each call has an O(d^3) set-up)
    rate = pilot_chains * N_TOTAL / max(el, 1e-6)
    n_chains = int(min(4096, max(cores, round(target_seconds * rate / N_TOTAL))))
    for _ in range(warmup):
        run(n_chains, N_BURNIN, N_KEEP)
    els = [run(n_chains, N_BURNIN, N_KEEP) for _ in range(steps)]
    per_step = sum(els) / len(els)
    return dict(value=n_chains * N_TOTAL / per_step, cores=cores, kind=kind, seconds_per_step=per_step,
                sample="%d chains x %d draws per step (same d, L, eps, seeds %d+c), %d host threads" % (n_chains, N_TOTAL, SEED, cores))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_throughput(steps=args.steps, warmup=min(args.warmup, 1), target_seconds=6.0)
    line = {
        "impl": "reference", "metric": "HMC draws/sec (chains x iters, d=128)", "value": r["value"], "unit": "draws/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": r["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference has no GPU path: host OpenMP loop over chains, one mcmc::hmc call per chain"},
        "cpu_baseline": {"value": r["value"], "unit": "draws/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "draws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- the other BASELINE configs, chains sharded contiguously over the ranks --------------------------------
def c3_problem():
    import numpy as np

    rng = np.random.default_rng(7)
    d, n = 1024, 4096
    X = rng.normal(size=(n, d))
    beta = np.sin(np.arange(d))
    yv = X @ beta + 0.5 * rng.normal(size=n)
    A = X.T @ X / 0.25 + np.eye(d) / 100.0
    A = (A + A.T) / 2
    b = X.T @ yv / 0.25
    eps = 0.5 / np.sqrt(np.linalg.eigvalsh(A).max())
    return d, np.concatenate([A.ravel(), b]), eps, np.linalg.solve(A, b)


def c4_problem():
    import numpy as np

    rng = np.random.default_rng(11)
    d = 256
    q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    lam = np.logspace(0, 3, d)
    P = (q / lam) @ q.T
    return d, ((P + P.T) / 2).ravel()


def run_configs(which, rank, world, local_rank, allmax, fp64_peak, hbm_peak, with_cpu, launches):
    """Each config: 1 warm-up + timed launches; kernel time = CUDA events inside the library, max over ranks."""
    import numpy as np
    import torch

    import mcmc_b200
    from mcmc_b200 import api
    from mcmc_b200.dist import chain_shard

    out = {}
    stream = torch.cuda.current_stream().cuda_stream
    dev = torch.device("cuda", local_rank)
    ol = ref = None
    if with_cpu:
        ol, ref, _ = ref_lib()
    cores = usable_cpus()

    def timed(fn, reps=2):
        fn()
        ms = []
        for _ in range(reps):
            r = fn()
            ms.append(r["kernel_ms"])
            launches[0] += r["kernel_launches"]
        return allmax(sum(ms) / len(ms)), r

    if "c3" in which:
        d, td, eps, mode = c3_problem()
        Ctot, nb, nk = 16384, 20, 100
        first, cnt = chain_shard(Ctot, rank, world)
        g = np.random.default_rng(1000 + rank)
        x0 = torch.from_numpy(g.normal(size=(cnt, d)) * 0.01 + mode).to(dev)
        draws = torch.empty((cnt, nk, d), dtype=torch.float64, device=dev)
        ms, r = timed(lambda: mcmc_b200.mala(None, "linreg", target_data=td, step_size=eps, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX, seed=3,
                                             initial_dev_ptr=x0.data_ptr(), n_chains=cnt, n_dim=d, draws_dev_ptr=draws.data_ptr(), stream=stream,
                                             chain_offset=first, device=local_rank))
        flops = 2.0 * d * d * Ctot * (nb + nk + 1)   # one gradient (A theta: the chain-batched GEMM) per draw, all ranks
        ach = flops / (ms * 1e-3) / 1e12 / world     # per GPU
        out["c3"] = {"workload": "C3: mcmc::mala, Bayesian linear regression d=1024 (sufficient statistics), %d chains total, %d+%d draws" % (Ctot, nb, nk),
                     "kernel_ms": ms, "launches_per_run": r["kernel_launches"], "value": Ctot * (nb + nk) / (ms * 1e-3), "unit": "draws/s",
                     "accept_rate": float(r["n_accept"].mean()) / nk,
                     "roofline": {"bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                                  "note": "fp64: 2 d^2 flop per gradient per chain; tcgen05 has no f64 kind, the GEMM is DMMA (mma.sync m8n8k4.f64) whose peak equals the DFMA peak measured in-process"}}
        del draws, x0
        if ref is not None:
            st = ol.Settings(n_burnin=0, n_keep=2, step_size=eps)
            x0c = g.normal(size=(min(cores, 8), d)) * 0.01 + mode
            t = ref.run_chains(ol.MALA, ol.TGT_LINREG, td, x0c, st, 3, n_threads=cores, keep_draws=False)[2]
            out["c3"]["cpu"] = {"value": x0c.shape[0] * 2 / t, "unit": "draws/s", "cores": min(cores, 8), "kind": "reference",
                                "sample": "%d chains x 2 draws (two O(d^3) dmvnorm factorisations per draw, SURVEY Q11)" % x0c.shape[0]}
    if "c4" in which:
        d, P = c4_problem()
        Ctot, nb, nk = 4096, 200, 200
        first, cnt = chain_shard(Ctot, rank, world)
        g = np.random.default_rng(2000)
        x0 = g.normal(size=(Ctot, d))[first:first + cnt]
        draws = torch.empty((cnt, nk, d), dtype=torch.float64, device=dev)
        x0d = torch.from_numpy(np.ascontiguousarray(x0)).to(dev)
        ms, r = timed(lambda: mcmc_b200.nuts(None, "dense_gauss", target_data=P, n_burnin=nb, n_keep=nk, n_adapt_draws=nb, rng_mode=api.RNG_PHILOX, seed=5,
                                             initial_dev_ptr=x0d.data_ptr(), n_chains=cnt, n_dim=d, draws_dev_ptr=draws.data_ptr(), stream=stream,
                                             chain_offset=first, device=local_rank), reps=1)
        nlf = torch.tensor([float(r["n_leapfrog"].sum())], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(nlf)
        nlf = float(nlf.item())
        ach = 2.0 * d * d * nlf / (ms * 1e-3) / 1e12 / world
        out["c4"] = {"workload": "C4: mcmc::nuts, dense Gaussian d=256 (cond 1e3), %d chains total sharded over %d GPU(s), %d adaptive + %d kept draws" % (Ctot, world, nb, nk),
                     "kernel_ms": ms, "value": Ctot * (nb + nk) / (ms * 1e-3), "unit": "draws/s", "leapfrogs_per_draw": nlf / (Ctot * (nb + nk)),
                     "leapfrogs_per_s": nlf / (ms * 1e-3), "step_size_mean": float(r["step_size"].mean()),
                     "roofline": {"bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                                  "note": "fp64: 2 d^2 flop (the Sigma^-1 x product) per distinct leapfrog state; per GPU; chain-batched rounds "
                                          "(csrc/nuts_batched.cu: one DMMA GEMM + one step-kernel launch per round and chain group) when the "
                                          "shard has >= 512 chains, else the persistent cooperative kernel (csrc/nuts.cu)"},
                     "kernel_launches": int(r["kernel_launches"])}
        del draws, x0d
        if ref is not None:
            st = ol.Settings(n_burnin=20, n_keep=10, n_adapt_draws=20)
            x0c = g.normal(size=(cores, d))
            t = ref.run_chains(ol.NUTS, ol.TGT_DENSE_GAUSS, P, x0c, st, 5, n_threads=cores, keep_draws=False)[2]
            out["c4"]["cpu"] = {"value": cores * 30 / t, "unit": "draws/s", "cores": cores, "kind": "reference", "sample": "%d chains x 30 draws" % cores}
    if "c5" in which:
        d, Ctot, nb, nk = 64, 2048, 2, 6
        first, cnt = chain_shard(Ctot, rank, world)
        g = np.random.default_rng(5)
        x0 = g.normal(size=(Ctot, d)) * 0.6
        x0[:, 0] = g.uniform(-0.5, 0.8, size=Ctot)
        x0s = np.ascontiguousarray(x0[first:first + cnt])
        ms, r = timed(lambda: mcmc_b200.rmhmc(x0s, "funnel", n_leap_steps=5, step_size=0.01, n_fp_steps=5, n_burnin=nb, n_keep=nk, rng_mode=api.RNG_PHILOX,
                                              seed=5, metric_id=2, chain_offset=first, device=local_rank, stream=stream), reps=1)
        L, nfp = 5, 5
        flop_draw = L * (nfp + 1) * 2.0 * d ** 3 + 2 * d ** 3 / 3.0   # the LU inverses (2 d^3 each) + the two Cholesky factorisations per draw
        ach = flop_draw * Ctot * (nb + nk) / (ms * 1e-3) / 1e12 / world
        out["c5"] = {"workload": "C5: mcmc::rmhmc, Neal's funnel d=64, SoftAbs metric (alpha=1e6), %d chains total, L=5, n_fp=5, eps=0.01, %d+%d draws" % (Ctot, nb, nk),
                     "kernel_ms": ms, "ms_per_draw": ms / (nb + nk), "value": Ctot * (nb + nk) / (ms * 1e-3), "unit": "draws/s",
                     "accept_rate": float(r["n_accept"].mean()) / nk, "finite_chains": float(np.isfinite(r["draws"]).all(axis=(1, 2)).mean()),
                     "roofline": {"bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                                  "note": "fp64 flops of the dense inverses/factorisations the reference algorithm prescribes per draw (L (n_fp+1) LU inverses of d x d); per GPU"}}
        if ref is not None:
            st = ol.Settings(n_burnin=0, n_keep=2, n_leap_steps=5, step_size=0.01, n_fp_steps=5, metric_id=2)
            x0c = x0[:cores]
            t = ref.run_chains(ol.RMHMC, ol.TGT_FUNNEL, None, x0c, st, 5, n_threads=cores, keep_draws=False)[2]
            out["c5"]["cpu"] = {"value": cores * 2 / t, "unit": "draws/s", "cores": cores, "kind": "reference", "sample": "%d chains x 2 draws" % cores}
    if "sweep" in which:
        sw = []
        for d in (32, 128, 512, 2048):
            Ctot, nb, nk = 4096, 100, 200
            first, cnt = chain_shard(Ctot, rank, world)
            x0 = torch.from_numpy(initial_vals(first, cnt, d)).to(dev)
            draws = torch.empty((cnt, nk, d), dtype=torch.float64, device=dev)
            ms, r = timed(lambda: mcmc_b200.hmc(None, "iso_gauss", n_leap_steps=10, step_size=0.1 * (128 / d) ** 0.25, n_burnin=nb, n_keep=nk,
                                                rng_mode=api.RNG_PHILOX, seed=1, initial_dev_ptr=x0.data_ptr(), n_chains=cnt, n_dim=d,
                                                draws_dev_ptr=draws.data_ptr(), stream=stream, chain_offset=first, device=local_rank))
            ach = Ctot * (nb + nk) * 2 * d * 8 / (ms * 1e-3) / 1e9 / world
            sw.append({"d": d, "chains_total": Ctot, "kernel_ms": ms, "value": Ctot * (nb + nk) / (ms * 1e-3), "unit": "draws/s",
                       "accept_rate": float(r["n_accept"].mean()) / nk,
                       "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak}})
            del draws, x0
        out["sweep"] = {"workload": "C5 dim sweep: mcmc::hmc iso-Gaussian, 4096 chains total, L=10, 100+200 draws, eps=0.1 (128/d)^(1/4)", "points": sw}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="strong,gather,c3,c4,c5,sweep", help="extra legs to run (comma list, '' = none)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)
    extras = set(x for x in args.configs.split(",") if x)

    import numpy as np
    import torch

    import mcmc_b200
    from mcmc_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank, world) if world > 1 else "single rank: not bound"
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dev = torch.device("cuda", local_rank)

    def allmax(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    C = CHAINS_PER_GPU
    first_chain = rank * C
    x0_host = torch.from_numpy(initial_vals(first_chain, C)).pin_memory()
    x0_dev = x0_host.to(dev)
    draws_dev = torch.empty((C, N_KEEP, D), dtype=torch.float64, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    common = dict(n_leap_steps=LEAP, step_size=EPS, n_burnin=N_BURNIN, n_keep=N_KEEP, rng_mode=api.RNG_PHILOX, seed=SEED,
                  arith=api.ARITH_FAST, device=local_rank, stream=stream)
    launches = [0]
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)

    def timed_steps(step_fn, steps, warmup):
        """W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream,
        max over ranks.  Returns (seconds per step, mean kernel ms (max over ranks), last result)."""
        for _ in range(warmup):
            step_fn()
        kms = []
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            r = step_fn()
            kms.append(r["kernel_ms"])
            launches[0] += r["kernel_launches"]
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        return allmax(ev0.elapsed_time(ev1) * 1e-3) / steps, allmax(sum(kms) / len(kms)), r, wall / steps

    def step_device(n_chains=C, first=first_chain, x0=x0_dev, out=draws_dev):
        flush.zero_()  # write a buffer larger than L2 between timed iterations
        return mcmc_b200.hmc(None, "iso_gauss", initial_dev_ptr=x0.data_ptr(), n_chains=n_chains, n_dim=D,
                             draws_dev_ptr=out.data_ptr(), chain_offset=first, **common)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    s_per_step, kernel_avg_ms, r, wall_step = timed_steps(step_device, args.steps, args.warmup)
    acc_rate = float(r["n_accept"].mean()) / N_KEEP
    ms_per_step = s_per_step * 1e3
    draws_per_step_all = world * C * N_TOTAL
    value = draws_per_step_all / s_per_step

    # ---- strong scaling: the same 4096 chains split over the ranks (north_star's target line) ------------------------
    strong = None
    if "strong" in extras:
        Cs = CHAINS_PER_GPU // world
        if world == 1:
            strong = {"chains_total": CHAINS_PER_GPU, "chains_per_gpu": Cs, "value": value, "ms_per_step": ms_per_step, "kernel_ms": kernel_avg_ms,
                      "note": "N=1: identical to the weak-scaling step"}
        else:
            xs = x0_dev[:Cs].clone()
            xs.copy_(torch.from_numpy(initial_vals(rank * Cs, Cs)))
            s2, k2, _, _ = timed_steps(lambda: step_device(Cs, rank * Cs, xs, draws_dev), args.steps, args.warmup)
            strong = {"chains_total": CHAINS_PER_GPU, "chains_per_gpu": Cs, "value": CHAINS_PER_GPU * N_TOTAL / s2, "ms_per_step": s2 * 1e3,
                      "kernel_ms": k2, "kernel_value": CHAINS_PER_GPU * N_TOTAL / (k2 * 1e-3),
                      "note": "value: K steps incl. the 256 MiB L2 flush and launch overhead of every step; kernel_value: CUDA-event kernel time only "
                              "(max over ranks).  %d chains/GPU = %d warps on 592 SM sub-partitions: latency-bound, see DESIGN.md §7" % (Cs, Cs)}

    # ---- weak-scaling steps that also assemble draws_out on every rank (library NCCL all-gather inside the timed region)
    gather = None
    value_with_gather = value if world == 1 else None
    if dist is not None and "gather" in extras:
        try:
            def xchg(idb):
                t = torch.zeros(128, dtype=torch.uint8, device=dev)
                if idb is not None:
                    t.copy_(torch.frombuffer(bytearray(idb), dtype=torch.uint8))
                dist.broadcast(t, 0)
                return bytes(t.cpu().numpy().tobytes())

            comm = api.Comm(world, rank, local_rank, xchg)
            full = torch.empty((world * C, N_KEEP, D), dtype=torch.float64, device=dev)
            cpr = [C] * world

            def step_gather():
                rr = step_device()
                comm.allgather_draws(draws_dev.data_ptr(), cpr, N_KEEP, D, full.data_ptr(), stream)
                return rr

            gs = max(2, min(args.steps, 3))
            s3, _, _, _ = timed_steps(step_gather, gs, 1)
            # gather alone
            barrier(); ev0.record()
            comm.allgather_draws(draws_dev.data_ptr(), cpr, N_KEEP, D, full.data_ptr(), stream)
            ev1.record(); barrier()
            gms = allmax(ev0.elapsed_time(ev1))
            ok = bool(torch.equal(full[rank * C:(rank + 1) * C, -1], draws_dev[:, -1]))
            value_with_gather = draws_per_step_all / s3
            recv = (world - 1) * C * N_KEEP * D * 8
            gather = {"ms": gms, "ms_per_step_with_gather": s3 * 1e3, "bytes_received_per_rank": recv, "recv_gbs_per_rank": recv / (gms * 1e-3) / 1e9,
                      "own_block_intact": ok, "steps": gs,
                      "note": "mcmcb200_allgather_draws (NCCL bound at run time, NVLink): every rank ends with all %d chains' draws (%.1f GB); "
                              "NVLink-bandwidth-bound, nothing to overlap with a %.1f ms kernel" % (world * C, world * C * N_KEEP * D * 8 / 1e9, kernel_avg_ms)}
            comm.destroy()
            del full
        except Exception as e:  # noqa: BLE001
            gather = {"error": str(e)[:300]}

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    e2e = None
    d2h = None
    if not args.no_e2e:
        draws_host = torch.empty((C, N_KEEP, D), dtype=torch.float64).pin_memory()
        draws_np = draws_host.numpy()
        x0_np = x0_host.numpy()
        e2e_steps = max(3, min(args.steps, 5))
        # what the host allows: a bare cudaMemcpy D2H of the same bytes, all ranks at once
        draws_host.copy_(draws_dev, non_blocking=True); barrier()
        ev0.record()
        for _ in range(2):
            draws_host.copy_(draws_dev, non_blocking=True)
        ev1.record(); barrier()
        d2h_ms = allmax(ev0.elapsed_time(ev1)) / 2
        d2h = {"ms": d2h_ms, "gbs_per_rank": draws_np.nbytes / (d2h_ms * 1e-3) / 1e9, "gbs_aggregate": world * draws_np.nbytes / (d2h_ms * 1e-3) / 1e9,
               "note": "bare pinned cudaMemcpyAsync D2H of draws_out (%.2f GB per rank), all %d rank(s) at once" % (draws_np.nbytes / 1e9, world)}

        def step_host():
            return mcmc_b200.hmc(x0_np, "iso_gauss", draws_out=draws_np, chain_offset=first_chain, **common)

        s4, _, rh, _ = timed_steps(step_host, e2e_steps, 1)
        e2e_ms = s4 * 1e3
        assert np.isfinite(draws_np[0, -1]).all() and int(rh["n_accept"].sum()) > 0
        e2e = {"value": draws_per_step_all / s4, "unit": "draws/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(x0_np.nbytes), "d2h_bytes_per_step": int(draws_np.nbytes + 8 * C),
               "steps": e2e_steps, "host_memory": "pinned", "frac_of_d2h_ceiling": d2h_ms / e2e_ms}
        del draws_host, draws_np

    # ---- same call chain, but the caller only needs posterior summaries: draws stay in HBM and are reduced there ------
    e2e_summary = None
    if not args.no_e2e:
        def step_summary():
            x0_dev.copy_(x0_host, non_blocking=True)   # H2D of this step's inputs from pinned memory
            rr = mcmc_b200.hmc(None, "iso_gauss", initial_dev_ptr=x0_dev.data_ptr(), n_chains=C, n_dim=D,
                               draws_dev_ptr=draws_dev.data_ptr(), chain_offset=first_chain, **common)
            step_summary.sm = api.summarize(draws_dev_ptr=draws_dev.data_ptr(), n_chains=C, n_keep=N_KEEP, n_dim=D, device=local_rank, stream=stream)
            rr["kernel_launches"] += 2
            return rr

        s5, _, _, _ = timed_steps(step_summary, e2e_steps, 1)
        sm = step_summary.sm
        assert abs(float(sm["var"].mean()) - 1.0) < 0.05 and float(sm["rhat"].max()) < 1.1
        e2e_summary = {"value": draws_per_step_all / s5, "unit": "draws/s", "ms_per_step": s5 * 1e3,
                       "h2d_bytes_per_step": int(x0_host.numpy().nbytes), "d2h_bytes_per_step": int(3 * D * 8 + 8 * C),
                       "summary_kernel_ms": sm["kernel_ms"],
                       "note": "draws_out stays in HBM; mcmcb200_summarize_draws returns mean/var/R-hat per element (not the reference's output format: informational)"}

    clocks = sampler.stop() if sampler else None
    del draws_dev

    # ---- the reference-shaped C++ call, rank 0 ------------------------------------------------------------------------
    e2e_cpp = None
    exe = os.path.join(ROOT, "mcmc_b200", "bin", "e2e_cpp")
    if not args.no_e2e and rank == 0 and os.path.exists(exe):
        try:
            p = subprocess.run([exe, str(C), str(N_BURNIN), str(N_KEEP), str(local_rank), "3"], capture_output=True, text=True, timeout=300)
            e2e_cpp = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": (p.stderr or p.stdout)[-300:]}
        except Exception as e:  # noqa: BLE001
            e2e_cpp = {"error": str(e)[:300]}
    barrier()

    # ---- other BASELINE configs ---------------------------------------------------------------------------------------
    peak, peak_src = peaks()
    configs = None
    which = extras & {"c3", "c4", "c5", "sweep"}
    fp64_peak = None
    if which:
        try:
            fp64_peak = allmax(api.fp64_peak(local_rank))
        except Exception:  # noqa: BLE001
            fp64_peak = None
        try:
            configs = run_configs(which, rank, world, local_rank, allmax, fp64_peak, peak, world == 1 and not args.no_cpu_baseline, launches)
            configs["fp64_peak_tflops"] = {"value": fp64_peak, "how": "dependent-free DFMA loop on every SM, CUDA events, in this process (mcmcb200_fp64_peak)"}
        except Exception as e:  # noqa: BLE001
            configs = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    if rank == 0:
        alg_bytes = ALG_BYTES_PER_DRAW * C * N_TOTAL  # per launch (one GPU)
        achieved = alg_bytes / (kernel_avg_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get("hmc_kernel_dram_bytes_per_launch")
            traffic_src = tj.get("source")
        line = {
            "metric": "HMC draws/sec (chains x iters, d=128)", "value": value, "unit": "draws/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_total": world * C, "leapfrog_steps_per_s": value * LEAP,
                       "rng": "Philox4x32-10 in-kernel", "arith": "fast (FMA)", "parallelism": "chains sharded, %d rank(s)" % world, "host_binding_rank0": numa,
                       "l2": "256 MiB buffer (2x L2) written between timed iterations; each step also writes 4.19 GB of draws (33x L2)",
                       "accept_rate": acc_rate, "wall_ms_per_step_rank0": wall_step * 1e3},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "hmc_pipe_kernel<IsoGauss,EPL=4,L=10>",
                         "kernel_ms": kernel_avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "2*d*8 B per transition x chains x draws per launch; fp64 instruction dispatch is the co-roof (DESIGN.md §4.1); "
                                 "traffic = dram__bytes_read+write of one ncu --set full capture of this kernel (file named in traffic_source), not re-measured per run"},
            "e2e": e2e, "gpu_launches": launches[0], "clocks": clocks,
            "value_with_gather": value_with_gather,
        }
        for k, v in (("strong", strong), ("allgather", gather), ("d2h_ceiling", d2h), ("e2e_cpp", e2e_cpp), ("e2e_summary", e2e_summary), ("configs", configs)):
            if v is not None:
                line[k] = v
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_throughput(steps=1, warmup=0, target_seconds=10.0)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "draws/s", "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"]}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
