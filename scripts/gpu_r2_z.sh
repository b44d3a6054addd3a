#!/bin/bash
mkdir -p gpurun_out
SG_HOST_TRACE=1 timeout 300 python scripts/probe_e2e.py 2>&1 | tail -8
