#!/bin/sh
# Builds the CPU oracle (test infrastructure) into oracle/_build/liboracle.so.
# -ffp-contract=off: rustc never fuses a*b+c; every product must round.
set -e
cd "$(dirname "$0")"
mkdir -p _build
gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -fvisibility=hidden -shared -fPIC \
    -o _build/liboracle.so shodh_oracle.c vamana_oracle.c pq_oracle.c spann_oracle.c -lm
echo "built oracle/_build/liboracle.so"
