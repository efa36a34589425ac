#!/usr/bin/env bash
# Round 2, GPU session I: whole GPU suite after the het tiling fix + in-sweep ADE tests.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/i_pytest_gpu.log 2>&1; tail -12 $O/i_pytest_gpu.log
