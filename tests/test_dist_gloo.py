"""CPU, world_size 2, gloo: the N>1 host path -- shard bounds, fitness all-gather in global-index
order, identical selection on every rank.  The per-shard evaluation here is the CPU oracle (this
test is about the exchange step; the GPU kernels are covered by the -m gpu tests)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from sgmm_b200 import synthetic
    from sgmm_b200.dist import all_gather_packed, all_gather_slices, shard_bounds
    P = 37                                            # not divisible by 2: last shard is short
    bundle = synthetic.synthetic_bundle(1, first_day=3)
    stats = synthetic.train_stats_of(bundle)
    z1, z2 = oracle.normalise(bundle, stats)
    bz = (z1, z2) + bundle[2:]
    master, _ = synthetic.policy_like_genomes(1, seed=5, out_scale=6.0, out_bias=(0.1, 0.1))
    first, count, stride = shard_bounds(P, world, rank)
    fit, trd = oracle.rollout_population(bz, 1e-4, 0.001, 0.0, master=master, sigma=0.05, seed=11, generation=2,
                                         first_index=first, count=count, nthreads=2)
    fit_all = all_gather_slices(torch.from_numpy(fit), P, stride, float("-inf"))
    trd_all = all_gather_slices(torch.from_numpy(trd), P, stride, 0)
    # the generation's real exchange: ONE all-gather of the packed (fitness | trades) block of every rank
    fit_p, trd_p = all_gather_packed(fit, trd, P, stride)
    assert torch.equal(fit_p, fit_all) and torch.equal(trd_p, trd_all)
    best = oracle.argmax(fit_all.numpy())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), fit=fit_all.numpy(), trd=trd_all.numpy(), best=best)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_selection(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["fit"], r1["fit"]) and np.array_equal(r0["trd"], r1["trd"])
    assert int(r0["best"]) == int(r1["best"])
    # identical to the unsharded population
    from oracle import oracle
    from sgmm_b200 import synthetic
    bundle = synthetic.synthetic_bundle(1, first_day=3)
    stats = synthetic.train_stats_of(bundle)
    z1, z2 = oracle.normalise(bundle, stats)
    master, _ = synthetic.policy_like_genomes(1, seed=5, out_scale=6.0, out_bias=(0.1, 0.1))
    fit, trd = oracle.rollout_population((z1, z2) + bundle[2:], 1e-4, 0.001, 0.0, master=master, sigma=0.05,
                                         seed=11, generation=2, first_index=0, count=37, nthreads=2)
    assert np.array_equal(fit, r0["fit"]) and np.array_equal(trd, r0["trd"])
    assert int(r0["best"]) == int(np.argmax(fit))


def test_block_layout_round_trip_and_shard_bounds():
    """Host twin of the library's rank-blocked gather buffer (include/sgmm.h, sgmm_ga_buffers)."""
    import torch
    from sgmm_b200.dist import block_layout, pack_block, shard_bounds, unpack_blocks
    for P, world in ((37, 2), (5, 4), (1001, 8), (8, 8), (3, 8)):
        bounds = [shard_bounds(P, world, r) for r in range(world)]
        stride = bounds[0][2]
        assert sum(c for _, c, _ in bounds) == P and all(s == stride for _, _, s in bounds)
        assert all(f == min(r * stride, P) for r, (f, _, _) in enumerate(bounds))
        nbytes, off = block_layout(stride)
        assert nbytes % 16 == 0 and off == stride * 8 and nbytes >= stride * 12
        fit = torch.arange(P, dtype=torch.float64) * 0.5 - 3
        trd = torch.arange(P, dtype=torch.int32) * 7
        buf = torch.cat([pack_block(fit[f:f + c], trd[f:f + c], stride) for f, c, _ in bounds])
        f2, t2 = unpack_blocks(buf, world, stride, P)
        assert torch.equal(f2, fit) and torch.equal(t2, trd)
