#!/bin/sh
# Builds the CPU oracle (test infrastructure) -> oracle/_build/liborc.so
set -e
cd "$(dirname "$0")"
mkdir -p _build
gcc -O2 -std=c99 -ffp-contract=off -fPIC -shared -fvisibility=hidden -o _build/liborc.so pcdet_oracle.c -lm
echo "built oracle/_build/liborc.so"
