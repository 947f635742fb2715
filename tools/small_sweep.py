"""Time the exact rollout for small populations; run once with SGMM_SMALL_POP_MAX=0 (sequential kernel) and once with
SGMM_SMALL_POP_MAX=100000 (policy-table path, sgmm_one.cu) to find the break-even.  args: tag"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sgmm_b200
from sgmm_b200 import synthetic
tag = sys.argv[1]
out = {}
for days in (1, 4, 12, 60):
    bundle = synthetic.synthetic_bundle(days)
    bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
    for P in (1, 8, 25, 50, 100, 148, 200, 256, 288, 320, 400, 500):
        m, genomes = synthetic.policy_like_genomes(P, seed=0)
        g = torch.from_numpy(genomes).cuda(); md = torch.from_numpy(m).cuda()
        for kind, run in (("explicit", lambda: sgmm_b200.rollout_population(bun, g, phi=1e-4)),
                          ("seeded", lambda: sgmm_b200.rollout_seeded(bun, md, count=P, sigma=0.05, seed=1, generation=0, phi=1e-4))):
            for _ in range(2):
                f, t = run()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()                    # graph replay: device time without the host's launch overhead
            with torch.cuda.graph(gr):
                for _ in range(5):
                    f, t = run()
            gr.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            out[f"{bun.T}.{P}.{kind}"] = (ms, f.double().sum().item(), int(t.sum().item()))
            print(f"[{tag}] T={bun.T:5d} P={P:4d} {kind:8s} {ms:8.4f} ms  {P * bun.T / ms / 1e6:7.3f} G env-steps/s  checksum {f.double().sum().item():.9f}", flush=True)
json.dump(out, open(f"gpurun_out/small_sweep_{tag}.json", "w"))
