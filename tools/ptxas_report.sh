#!/usr/bin/env bash
# Rebuild libfdtd_b200.so with -Xptxas -v and list registers / spills per kernel (no GPU needed).
cd "$(dirname "$0")/../prismo_b200/csrc" || exit 1
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function \
     -shared -Xptxas -v ${NVCC_EXTRA:-} -o ../libfdtd_b200.so fdtd_engine.cu > /tmp/ptxas.log 2>&1 || { cat /tmp/ptxas.log | tail -30; exit 1; }
python - "$@" <<'PY'
import re, sys, subprocess
pat = sys.argv[1] if len(sys.argv) > 1 else None
t = open('/tmp/ptxas.log').read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t):
    n = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
    if (pat and re.search(pat, n)) or (not pat and int(m.group(2)) > 0):
        print(f"{n[:80]:80s} stack {m.group(2):>4s} spill st {m.group(3):>4s} ld {m.group(4):>4s} regs {m.group(5)}")
PY
