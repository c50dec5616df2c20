#!/bin/bash
# cli_trace.sh -- where the wall clock of the host tools goes: one C4-shaped file, ZDW_HOST_TIMING=1 (run on the GPU box)
set -e
cd "$(dirname "$0")/.."
D=$(mktemp -d -p /dev/shm zdwtrace_XXXX)
python - "$D" "${1:-8}" <<'PY'
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
d, blocks = sys.argv[1], int(sys.argv[2])
s = bench.Synth(); cap = s.cap_for(bench.ROWS_PER_BLOCK); buf = (C.c_uint8 * cap)()
with open(d + "/x.sql", "wb") as f:
    for b in range(blocks):
        n = s.block_into(b, bench.ROWS_PER_BLOCK, C.addressof(buf), cap); f.write(memoryview(buf)[:n])
open(d + "/x.desc.sql", "wb").write(s.desc)
PY
export PATH=$PWD/oracle/_ref/nocomp:$PATH ZDW_HOST_TIMING=1
cd $D
for extra in "" "--lanes-per-gpu=1" "--lanes-per-gpu=4"; do
  echo "== convertDWfile -q $extra"; rm -f x.zdw.gz
  T0=$(date +%s.%N); $OLDPWD/zdw_b200/bin/convertDWfile -q $extra x.sql 2>&1 | tail -12; python3 -c "import time,sys; print('wall %.3f s' % (time.time() - float(sys.argv[1])))" $T0
done
mkdir -p out
for extra in "" "--lanes-per-gpu=1" "--lanes-per-gpu=4"; do
  echo "== unconvertDWfile -q $extra"; rm -f out/*
  T0=$(date +%s.%N); $OLDPWD/zdw_b200/bin/unconvertDWfile -q $extra -d out x.zdw.gz 2>&1 | tail -30; python3 -c "import time,sys; print('wall %.3f s' % (time.time() - float(sys.argv[1])))" $T0
  cmp out/x.sql x.sql && echo identical
done
echo "== plain copies for scale"; T0=$(date +%s.%N); cp x.sql out/y.sql; python3 -c "import time,sys; print('cp wall %.3f s' % (time.time() - float(sys.argv[1])))" $T0; T0=$(date +%s.%N); cat x.sql > /dev/null; python3 -c "import time,sys; print('cat wall %.3f s' % (time.time() - float(sys.argv[1])))" $T0
cd /; rm -rf $D
