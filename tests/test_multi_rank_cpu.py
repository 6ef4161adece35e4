"""world_size-2 gloo test (CPU) of the N>1 path: chain sharding + all-gather assembly of draws_out.  The sampler on
each rank is the CPU oracle in Philox mode with the rank's global chain offset — exactly the arguments the CUDA
engine receives on a GPU box — so the test checks that sharded runs reproduce the single-process result chain for
chain (the kernels' side of the same property is tests/test_gpu_hmc.py::test_full_size_c2_properties)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol

N_CHAINS, D, WORLD = 7, 10, 2   # 7 chains over 2 ranks: uneven shards (4 + 3)


def _run_chains(first, count):
    orc = ol.Oracle()
    st = ol.Settings(n_burnin=2, n_keep=5, n_leap_steps=3, step_size=0.2)
    x0 = ol.c2_initial(count, D, first)
    out = np.zeros((count, 5, D))
    for c in range(count):
        out[c] = orc.run_chain(ol.HMC, ol.TGT_ISO_GAUSS, None, x0[c], st, seed=99, rng_mode=ol.RNG_PHILOX, chain_id=first + c,
                               sum_mode=ol.SUM_WARP)["draws"]
    return out


def _worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from mcmc_b200.dist import all_gather_draws, all_gather_summary, all_reduce_max, chain_shard

    first, count = chain_shard(N_CHAINS, rank, WORLD)
    local = torch.from_numpy(_run_chains(first, count))
    full = all_gather_draws(local, N_CHAINS)
    t = all_reduce_max(1.0 + rank, "cpu")
    # summaries of all ranks' chains from per-chain statistics only (what api.summarize(per_chain=True) returns on a GPU)
    m, v, rh = all_gather_summary(local.mean(dim=1), local.var(dim=1, unbiased=True), local.shape[1], N_CHAINS)
    q.put((rank, first, count, full.numpy(), t, m.numpy(), v.numpy(), rh.numpy()))
    dist.destroy_process_group()


def test_sharded_runs_assemble_to_single_process_result():
    from mcmc_b200.dist import chain_shard

    assert [chain_shard(7, r, 2) for r in range(2)] == [(0, 4), (4, 3)]
    assert [chain_shard(4096 * 8, r, 8) for r in (0, 7)] == [(0, 4096), (28672, 4096)]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(WORLD)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _run_chains(0, N_CHAINS)
    T = want.shape[1]
    cm, cv = want.mean(axis=1), want.var(axis=1, ddof=1)
    W, Bn = cv.mean(axis=0), cm.var(axis=0, ddof=1)
    for rank, first, count, full, tmax, m, v, rh in res:
        assert full.shape == (N_CHAINS, 5, D)
        assert np.array_equal(full, want)
        assert tmax == 2.0
        assert np.allclose(m, want.reshape(-1, D).mean(axis=0), rtol=1e-13, atol=1e-15)
        assert np.allclose(v, want.reshape(-1, D).var(axis=0, ddof=1), rtol=1e-12)
        assert np.allclose(rh, np.sqrt(((T - 1) / T * W + Bn) / W), rtol=1e-12)


def test_chain_shard_is_a_contiguous_balanced_partition():
    """mcmc_b200.dist.chain_shard for every (n_chains, world_size) a node can see: the shards tile [0, n_chains) in rank order,
    sizes differ by at most one, remainders go to the lowest ranks — so chain_offset (the global chain id the Philox counter and
    the MT seed are derived from) is the same whatever the number of ranks."""
    from mcmc_b200.dist import chain_shard

    for world in (1, 2, 3, 4, 7, 8, 16):
        for n in list(range(0, 40)) + [511, 512, 513, 2048, 4096, 16384, 16385]:
            nxt, sizes = 0, []
            for r in range(world):
                first, count = chain_shard(n, r, world)
                assert first == nxt and count >= 0
                nxt += count
                sizes.append(count)
            assert nxt == n and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
