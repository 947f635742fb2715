"""torchrun --nproc-per-node N tools/sharded_ga_check.py
Sharded GA (one process per GPU, NCCL all-gather of fitness slices) must reproduce the single-rank GA
bit for bit: children are regenerated from the counter-based stream, so sharding cannot change them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import sgmm_b200
from sgmm_b200 import synthetic
from sgmm_b200.dist import ShardedGA
from sgmm_b200.engine import DeviceGA

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
tb = synthetic.synthetic_bundle(3, first_day=80)
vb = synthetic.synthetic_bundle(1, first_day=83)
stats = synthetic.train_stats_of(tb)
train = sgmm_b200.Bundle.from_arrays(tb, stats, 0.001, device=local)
val = sgmm_b200.Bundle.from_arrays(vb, stats, 0.001, device=local)
master, _ = synthetic.policy_like_genomes(1, seed=31)
adv_master = (np.random.default_rng(6).standard_normal(1250) * 0.5).astype(np.float32)
POP, GENS = 1001, 6          # not divisible by the world size on purpose
kw = dict(pop_size=POP, sigma=0.05, phi=1e-4, fee_rate=0.0, use_arl=True, seed=99, max_generations=GENS, patience=2,
          device=local)
sh = ShardedGA(lambda shard: DeviceGA(master, adv_master, shard=shard, **kw), POP)
for _ in range(GENS):
    sh.generation(train, val)
h = sh.ga.history(GENS)
mm, adv, best = sh.ga.masters()
ref = DeviceGA(master, adv_master, **kw)
for _ in range(GENS):
    ref.generation(train, val)
hr = ref.history(GENS)
mmr, advr, bestr = ref.masters()
ok = all(np.array_equal(h[k].view(np.uint8), hr[k].view(np.uint8)) for k in h)
ok = ok and np.array_equal(mm, mmr) and np.array_equal(adv, advr) and np.array_equal(best, bestr)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded GA on {world} ranks == single-rank GA bit for bit: {bool(flag.item())};  val_f {h['val_f']}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
