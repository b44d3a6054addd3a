#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -4
CASES=256x10,256x120,128x10,1x100,4x10 NS=1250000 REPS=300 timeout 300 python scripts/probe_step.py 2>&1 | tail -5
