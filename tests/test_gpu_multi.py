"""GPU (-m gpu), needs >= 2 visible devices (skipped otherwise; run with `gpurun --gpus 2`): the population sharded over
ranks THROUGH THE REFERENCE-NAMED API -- DRLEngine.train with torch.distributed initialised (NCCL) -- must reproduce the
single-rank run bit for bit: children come from the counter-based stream, every rank all-gathers the same packed
(fitness, trades) blocks with ONE collective per generation and runs the identical argmax / tell / validate / select.
Replaces Pool(8).starmap of Env/drl_engine.py:91,115."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _data():
    from sgmm_b200 import synthetic
    tb = synthetic.synthetic_bundle(2, first_day=170)
    vb = synthetic.synthetic_bundle(1, first_day=172)
    return tb, vb, synthetic.train_stats_of(tb)


def _train(save_dir, use_arl, precision, device):
    import sgmm_b200
    tb, vb, stats = _data()
    torch.manual_seed(11)
    eng = sgmm_b200.DRLEngine(pop_size=101, phi=1e-4, tick_size=0.001, save_dir=save_dir, seed=5, use_arl=use_arl,
                              precision=precision, device=device, patience=2)
    policy, hist = eng.train(tb, vb, stats, generations=6, verbose=False)
    return policy.get_weights().numpy(), hist, eng


def _worker(rank, world, port, out_dir, use_arl, precision):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    w, hist, eng = _train(os.path.join(out_dir, f"ck{rank}"), use_arl, precision, rank)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), w=w, **{k: np.asarray(v, np.float64) for k, v in hist.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_arl,precision", [(False, None), (True, None), (True, "f16")])
def test_drl_engine_train_sharded_over_two_ranks_is_bit_identical(tmp_path, use_arl, precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible CUDA devices (gpurun --gpus 2)")
    world = 2
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path), use_arl, precision), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    w, hist, _ = _train(str(tmp_path / "single"), use_arl, precision, 0)
    for k in ("train_f", "val_f", "train_trades", "val_trades"):
        assert np.array_equal(r0[k], r1[k]), k                                   # every rank reports the same history
        assert np.array_equal(r0[k].view(np.uint64), np.asarray(hist[k], np.float64).view(np.uint64)), k
    assert np.array_equal(r0["w"], w) and np.array_equal(r1["w"], w)
    assert (tmp_path / "ck0" / "agent_best_val_0.0001.pth").exists()            # rank 0 writes the checkpoint
    assert not (tmp_path / "ck1" / "agent_best_val_0.0001.pth").exists()
