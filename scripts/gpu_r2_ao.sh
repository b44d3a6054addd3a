#!/bin/bash
# round 2, call ao: last check of the committed tree
mkdir -p gpurun_out
timeout 200 python scripts/probe_enc.py 2>&1 | tail -3
timeout 1500 python -m pytest tests -q --tb=short -m gpu -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/ao_tests.log; tail -5 gpurun_out/ao_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
