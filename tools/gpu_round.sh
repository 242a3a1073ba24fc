#!/usr/bin/env bash
# One gpurun call for a development round: isolated GPU tests for the listed files, then the
# default bench (with cpu_baseline). Usage: tools/gpu_round.sh tests/test_a.py [...]
set -u
mkdir -p gpurun_out
tools/gpu_isolated.sh "$@"
