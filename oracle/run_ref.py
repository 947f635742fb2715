"""Run the UNMODIFIED reference (staged under oracle/_ref by oracle/stage_ref.py) on inputs handed over in an .npz.

    python oracle/run_ref.py --in IN.npz --out OUT.npz [--mode single|pool] [--procs 8] [--torch-threads N]
                             [--steps K] [--warmup W]

Always a separate process: the reference's modules are called `Env` and `models`, the same names the drop-in shims of
this repository shadow, and its Pool workers are forked from whoever imports it.  Test / measurement infrastructure only
(tests/, bench.py's `--impl reference` and `cpu_baseline` legs); the product never runs this.

IN.npz   s1 s2 mid_next best_ask best_bid buy_max sell_min   the reference's 7-tuple bundle (pipeline/agent_trainer.py:75-77)
         s1_m s1_s s2_m s2_s                                 train_stats, dtypes preserved (agent_trainer.py:126-129)
         genomes [P,1250] f32, adv [P,G] f32 (optional), phi tick fee use_arl
OUT.npz  fitness f64[P], trades i64[P], seconds f64[K] (one entry per timed step), mode, procs, torch_threads

mode single : `[evaluate_individual(...) for each individual]` in this process (Env/drl_engine.py:9-67)
mode pool   : `Pool(processes=procs).starmap(partial(evaluate_individual, ...), zip(mm_pop, adv_pop))` exactly as
              DRLEngine.train does it (Env/drl_engine.py:91,104-115); the pool is created once, outside the timed steps,
              as in the reference (one pool per training run).  --torch-threads 0 leaves torch's default in the workers
              (the shipped behaviour); N > 0 calls torch.set_num_threads(N) in every worker (pool initializer).
"""
import argparse
import os
import sys
import time

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def _init_worker(n):
    import torch
    if n > 0:
        torch.set_num_threads(n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in", dest="inp", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--mode", default="single", choices=["single", "pool"])
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--torch-threads", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    a = ap.parse_args()
    sys.path.insert(0, HERE)
    import stage_ref
    if not stage_ref.staged(REF):
        raise SystemExit("oracle/_ref is missing or modified: run `python oracle/stage_ref.py` where /root/reference exists")
    sys.path.insert(0, REF)
    import numpy as np
    import torch
    from functools import partial
    from multiprocessing import Pool
    from Env.drl_engine import evaluate_individual        # the reference's own function, unmodified

    d = np.load(a.inp)
    bundle = tuple(d[k] for k in ("s1", "s2", "mid_next", "best_ask", "best_bid", "buy_max", "sell_min"))
    stats = {k: d[k][()] for k in ("s1_m", "s1_s", "s2_m", "s2_s")}
    use_arl = bool(d["use_arl"])
    mm_pop = [torch.from_numpy(g.copy()) for g in d["genomes"]]
    adv_pop = [torch.from_numpy(g.copy()) for g in d["adv"]] if (use_arl and "adv" in d.files) else [None] * len(mm_pop)
    kw = dict(bundle=bundle, phi=float(d["phi"]), tick_size=float(d["tick"]), fee_rate=float(d["fee"]),
              train_stats=stats, use_arl=use_arl)
    seconds, results = [], None
    if a.mode == "single":
        if a.torch_threads > 0:
            torch.set_num_threads(a.torch_threads)
        for i in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            results = [evaluate_individual(m, v, **kw) for m, v in zip(mm_pop, adv_pop)]
            if i >= a.warmup:
                seconds.append(time.perf_counter() - t0)
    else:
        eval_func = partial(evaluate_individual, **kw)                       # drl_engine.py:104-112
        with Pool(processes=a.procs, initializer=_init_worker, initargs=(a.torch_threads,)) as pool:   # :91
            for i in range(a.warmup + a.steps):
                t0 = time.perf_counter()
                results = pool.starmap(eval_func, zip(mm_pop, adv_pop))      # :115
                if i >= a.warmup:
                    seconds.append(time.perf_counter() - t0)
    np.savez(a.out, fitness=np.array([r[0] for r in results], np.float64),
             trades=np.array([r[1] for r in results], np.int64), seconds=np.array(seconds, np.float64),
             mode=a.mode, procs=a.procs, torch_threads=a.torch_threads,
             torch_version=torch.__version__, numpy_version=np.__version__)


if __name__ == "__main__":
    main()
