"""Time the exact rollout kernel under tuning knobs read from the environment at launch (SGMM_ROLLOUT_*).
args: P days  then a list of KEY=VAL,KEY=VAL settings (one timed run each); checks bit-identity against the first."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]); days = int(sys.argv[2]); settings = sys.argv[3:] or [""]
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
master, genomes = synthetic.policy_like_genomes(P, hidden=32, seed=0, out_scale=1.0)
g = torch.from_numpy(genomes).cuda()
ref = None
for s in settings:
    for k in list(os.environ):
        if k.startswith("SGMM_ROLLOUT_"):
            del os.environ[k]
    for kv in filter(None, s.split(",")):
        k, v = kv.split("=")
        os.environ["SGMM_ROLLOUT_" + k.upper()] = v
    for _ in range(2):
        f, t = sgmm_b200.rollout_population(bun, g, phi=1e-4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f, t = sgmm_b200.rollout_population(bun, g, phi=1e-4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    if ref is None:
        ref = (f.clone(), t.clone())
    same = torch.equal(f, ref[0]) and torch.equal(t, ref[1])
    print(f"P={P} T={bun.T} [{s or 'default'}]: {ms:.3f} ms  {P * bun.T / ms / 1e6:.2f} G env-steps/s  identical={same}", flush=True)
