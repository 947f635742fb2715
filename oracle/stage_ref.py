"""Stage the reference's own implementation of the hot path under oracle/_ref/ (git-ignored, travels to the GPU box).

    python oracle/stage_ref.py            # or: __graft_entry__.build(), which calls stage() when /root/reference exists

The reference is pure Python: there is nothing to compile, so "building oracle/_ref" = copying, byte for byte, the three
source files the path lives in (plus its caller) from where they lie under /root/reference,

    Env/market_env.py      FTPEnv.step / reset                      
    Env/drl_engine.py      evaluate_individual, DRLEngine            (a3-a5, a8, a11)
    models/model.py        TradingPolicy, AdversaryPolicy, NeuroEvolution (a4, a6, a7, a9, a10)
    pipeline/agent_trainer.py, pipeline/evaluator.py   the callers of the path (8b), for the drop-in proof

into oracle/_ref/ together with MANIFEST.json (sha256 of every file, so a test can prove they are unmodified).  Nothing
under oracle/_ref/ is ever committed (the reference's sources must not enter this repository's history) and the PRODUCT
never imports it: only bench.py's `--impl reference` / `cpu_baseline` legs and the `-m gpu` live-reference parity test
execute it, through oracle/run_ref.py in a separate process, as the thing compared AGAINST.
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ("Env/market_env.py", "Env/drl_engine.py", "models/model.py",
         # the path's CALLER, staged for the drop-in proof only (tests/test_gpu_dropin_pipeline.py runs it unchanged with
         # this repository's dropin/ shadowing Env.* and models.model); never imported by oracle/run_ref.py
         "pipeline/agent_trainer.py", "pipeline/evaluator.py")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(ref=REF, dst=DST):
    """Copy the files; returns the manifest dict.  Raises if the reference is not there."""
    manifest = {"source": ref, "files": {}}
    for rel in FILES:
        src = os.path.join(ref, rel)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(src, out)
        manifest["files"][rel] = _sha(out)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def staged(dst=DST):
    """True when oracle/_ref holds the three files and their hashes match the manifest (i.e. nobody edited them)."""
    mf = os.path.join(dst, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    try:
        files = json.load(open(mf))["files"]
    except (OSError, ValueError, KeyError):
        return False
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h for rel, h in files.items()) \
        and set(files) == set(FILES)


if __name__ == "__main__":
    m = stage()
    for rel, h in m["files"].items():
        print(f"{h[:16]}  oracle/_ref/{rel}")
