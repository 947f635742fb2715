#!/bin/bash
# rebuild libsgmm_b200.so from the repo root and print the tensor-core kernels' resource usage
cd "$(dirname "$0")/.." || exit 1
python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','deep-reinforcement-learning-based-signal-gated-market-making_b200/build.py');m=importlib.util.module_from_spec(spec);spec.loader.exec_module(m);m.build(force=True)" 2>&1 | tail -30
grep -n "${1:-tc32}" -A3 deep-reinforcement-learning-based-signal-gated-market-making_b200/csrc/build.log | grep -i "registers\|spill\|error" | head -6
if grep -q "error" deep-reinforcement-learning-based-signal-gated-market-making_b200/csrc/build.log; then echo "BUILD FAILED"; grep -n "error" deep-reinforcement-learning-based-signal-gated-market-making_b200/csrc/build.log | head -5; exit 1; fi
