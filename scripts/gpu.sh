#!/bin/bash
# usage: scripts/gpu.sh <timeout_s> <command...>   -- rebuilds every native piece first (a stale .so travels otherwise)
set -e
cd "$(dirname "$0")/.."
make -s -C ncrystal_b200/csrc 2>&1 | grep -v "^$" | grep -v "Remark" | tail -5 || true
make -s -C tests/hostsim 2>&1 | tail -5
make -s -C oracle oracle tools 2>&1 | tail -5
t=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
