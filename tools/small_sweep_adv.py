"""As small_sweep.py, with the adversary: run once with SGMM_SMALL_POP_MAX_ADV=0 and once with =100000.  args: tag"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sgmm_b200
from sgmm_b200 import synthetic
tag = sys.argv[1]
out = {}
for days in (1, 4, 12, 60):
    bundle = synthetic.synthetic_bundle(days)
    bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
    for P in (1, 8, 50, 100, 148, 200, 296):
        m, genomes = synthetic.policy_like_genomes(P, seed=0)
        am = (np.random.default_rng(2).standard_normal(1250) * 0.5).astype(np.float32)
        md = torch.from_numpy(m).cuda(); amd = torch.from_numpy(am).cuda()
        run = lambda: sgmm_b200.rollout_seeded(bun, md, count=P, sigma=0.05, seed=1, generation=0, adv_master=amd, phi=1e-4)
        for _ in range(2):
            f, t = run()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(5):
                f, t = run()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f"{bun.T}.{P}"] = (ms, f.double().sum().item(), int(t.sum().item()))
        print(f"[{tag}] T={bun.T:5d} P={P:4d} adv {ms:8.4f} ms  {P * bun.T / ms / 1e6:7.3f} G env-steps/s", flush=True)
json.dump(out, open(f"gpurun_out/small_sweep_adv_{tag}.json", "w"))
