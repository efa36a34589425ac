#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py yee > $O/e2_sanitize_memcheck_yee.log 2>&1; head -60 $O/e2_sanitize_memcheck_yee.log
