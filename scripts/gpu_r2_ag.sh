#!/bin/bash
# round 2, call ag: what the programmatic-dependent-launch chain is worth (encoder forward, search step)
mkdir -p gpurun_out
echo "== PDL (default)"; timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
echo "== plain launches"; SG_NO_PDL=1 timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
echo "== PDL, no graph"; SG_ENC_GRAPH=0 timeout 200 python scripts/probe_enc.py 2>&1 | tail -3 | head -1
echo "== plain, no graph"; SG_NO_PDL=1 SG_ENC_GRAPH=0 timeout 200 python scripts/probe_enc.py 2>&1 | tail -3 | head -1
NS=1250000 CASES=256x10 timeout 200 python scripts/probe_step.py 2>&1 | tail -1
SG_NO_PDL=1 NS=1250000 CASES=256x10 timeout 200 python scripts/probe_step.py 2>&1 | tail -1
