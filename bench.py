#!/usr/bin/env python
"""bench.py — the headline measurement of BASELINE.json: HMC draws/sec (chains x iterations, d=128).

A "step" is one pass of the hot path over one batch of synthetic input: the whole C2 job of SURVEY.md §8(d) —
mcmc::hmc semantics on log pi(x) = -|x|^2/2, d=128, 4096 chains PER GPU, L=10, eps=0.1, M=I, fp64,
n_burnin=100 + n_keep=1000 draws, x0[c][j]=sin(0.37c+0.11j), in-kernel Philox (seed 12345).  Chains shard
across ranks with no data-path collective (weak scaling: 4096 chains per GPU, global chain ids).

  value    draws/s of the whole job with inputs resident in HBM (device pointers through the C ABI)
  e2e      the same metric through the C ABI with HOST (pinned) buffers: H2D of x0 and D2H of draws_out and
           n_accept_draws inside the timed region
  roofline algorithmic bytes (2*d*8 B per transition) / CUDA-event kernel time vs the measured HBM peak
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


def initial_vals(first_chain, n_chains):
    import numpy as np

    c = np.arange(first_chain, first_chain + n_chains, dtype=np.float64)[:, None]
    j = np.arange(D, dtype=np.float64)[None, :]
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


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the host cores of the NUMA node its GPU hangs off (so the pinned staging buffers of the end-to-end
    leg are allocated there and the D2H stream does not cross the socket interconnect).  Best effort: any missing piece
    (sysfs entry, empty intersection with the cpuset) leaves the affinity untouched.  Returns a short description."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, allowed)
        return "numa node %d (%d cpus)" % (node, len(allowed))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

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
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dev = torch.device("cuda", local_rank)
    C = CHAINS_PER_GPU
    first_chain = rank * C
    x0_host = torch.from_numpy(initial_vals(first_chain, C)).pin_memory()
    x0_dev = x0_host.to(dev)
    draws_dev = torch.empty((C, N_KEEP, D), dtype=torch.float64, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    common = dict(n_leap_steps=LEAP, step_size=EPS, n_burnin=N_BURNIN, n_keep=N_KEEP, rng_mode=api.RNG_PHILOX, seed=SEED,
                  arith=api.ARITH_FAST, chain_offset=first_chain, device=local_rank, stream=stream)

    kernel_ms = []

    def step_device():
        flush.zero_()  # write a buffer larger than L2 between timed iterations
        r = mcmc_b200.hmc(None, "iso_gauss", initial_dev_ptr=x0_dev.data_ptr(), n_chains=C, n_dim=D,
                          draws_dev_ptr=draws_dev.data_ptr(), **common)
        kernel_ms.append(r["kernel_ms"])
        return r

    for _ in range(args.warmup):
        r = step_device()
    kernel_ms.clear()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        r = step_device()
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    elapsed = ev0.elapsed_time(ev1) * 1e-3  # device time (CUDA events on the launching stream) of exactly K steps
    acc_rate = float(r["n_accept"].mean()) / N_KEEP
    el = torch.tensor([elapsed], dtype=torch.float64, device=dev)
    km = torch.tensor([sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
    elapsed = float(el.item())
    kernel_avg_ms = float(km.item())
    ms_per_step = elapsed / args.steps * 1e3
    draws_per_step_all = world * C * N_TOTAL
    value = draws_per_step_all / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    e2e = None
    launches = args.steps
    if not args.no_e2e:
        draws_host = torch.empty((C, N_KEEP, D), dtype=torch.float64).pin_memory()
        draws_np = draws_host.numpy()
        x0_np = x0_host.numpy()
        e2e_steps = max(3, min(args.steps, 5))

        def step_host():
            return mcmc_b200.hmc(x0_np, "iso_gauss", draws_out=draws_np, **common)

        step_host()
        barrier()
        ev0.record()
        for _ in range(e2e_steps):
            rh = step_host()
        ev1.record()
        barrier()
        e2 = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(e2, op=dist.ReduceOp.MAX)
        e2e_ms = float(e2.item()) / e2e_steps * 1e3
        launches += e2e_steps
        assert np.isfinite(draws_np[0, -1]).all() and int(rh["n_accept"].sum()) > 0
        e2e = {"value": draws_per_step_all / (e2e_ms * 1e-3), "unit": "draws/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(x0_np.nbytes), "d2h_bytes_per_step": int(draws_np.nbytes + 8 * C),
               "steps": e2e_steps, "host_memory": "pinned"}
        del draws_host, draws_np

    # ---- same call chain, but the caller only needs posterior summaries: draws stay in HBM and are reduced there ------
    e2e_summary = None
    if not args.no_e2e:
        def step_summary():
            x0_dev.copy_(x0_host, non_blocking=True)   # H2D of this step's inputs from pinned memory
            mcmc_b200.hmc(None, "iso_gauss", initial_dev_ptr=x0_dev.data_ptr(), n_chains=C, n_dim=D,
                          draws_dev_ptr=draws_dev.data_ptr(), **common)
            return api.summarize(draws_dev_ptr=draws_dev.data_ptr(), n_chains=C, n_keep=N_KEEP, n_dim=D, device=local_rank, stream=stream)

        step_summary()
        barrier()
        ev0.record()
        for _ in range(e2e_steps):
            sm = step_summary()
        ev1.record()
        barrier()
        e3 = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(e3, op=dist.ReduceOp.MAX)
        es_ms = float(e3.item()) / e2e_steps * 1e3
        launches += 3 * e2e_steps   # hmc + the two reduction kernels per step
        assert abs(float(sm["var"].mean()) - 1.0) < 0.05 and float(sm["rhat"].max()) < 1.1
        e2e_summary = {"value": draws_per_step_all / (es_ms * 1e-3), "unit": "draws/s", "ms_per_step": es_ms,
                       "h2d_bytes_per_step": int(x0_np.nbytes), "d2h_bytes_per_step": int(3 * D * 8 + 8 * C),
                       "summary_kernel_ms": sm["kernel_ms"],
                       "note": "draws_out stays in HBM; mcmcb200_summarize_draws returns mean/var/R-hat per element (not the reference's output format: informational)"}

    clocks = sampler.stop() if sampler else None

    # ---- optional: assemble draws_out on every rank (north_star's all-gather), outside the timed region ---
    gather = None
    if dist is not None and os.environ.get("MCMCB200_BENCH_GATHER", "1") == "1":
        try:
            full = torch.empty((world * C, N_KEEP, D), dtype=torch.float64, device=dev)
            barrier()
            g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
            g0.record()
            dist.all_gather_into_tensor(full, draws_dev)
            g1.record()
            torch.cuda.synchronize()
            gm = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(gm, op=dist.ReduceOp.MAX)
            gather = {"ms": float(gm.item()), "bytes_per_rank_out": int(full.numel() * 8),
                      "note": "ncclAllGather of draws_out over NVLink, not inside the timed steps"}
            del full
        except Exception as e:  # noqa: BLE001
            gather = {"error": str(e)[:200]}

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = ALG_BYTES_PER_DRAW * C * N_TOTAL  # per launch (one GPU)
        achieved = alg_bytes / (kernel_avg_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("hmc_kernel_dram_bytes_per_launch")
        line = {
            "metric": "HMC draws/sec (chains x iters, d=128)", "value": value, "unit": "draws/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_total": world * C, "leapfrog_steps_per_s": value * LEAP,
                       "rng": "Philox4x32-10 in-kernel", "arith": "fast (FMA)", "parallelism": "chains sharded, %d rank(s)" % world, "host_binding_rank0": numa,
                       "l2": "256 MiB buffer (2x L2) written between timed iterations; each step also writes 4.19 GB of draws (33x L2)",
                       "accept_rate": acc_rate, "wall_ms_per_step_rank0": wall / args.steps * 1e3},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "hmc_pipe_kernel<IsoGauss,EPL=4,L=10>",
                         "kernel_ms": kernel_avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "2*d*8 B per transition x chains x draws per launch; fp64 instruction dispatch is the co-roof (DESIGN.md §4.1)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if e2e_summary is not None:
            line["e2e_summary"] = e2e_summary
        if gather is not None:
            line["allgather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_throughput(steps=1, warmup=0, target_seconds=12.0)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "draws/s", "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"]}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
