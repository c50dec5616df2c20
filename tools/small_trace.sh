set -e
cd /root/repo
D=$(mktemp -d -p /dev/shm zdwtrace_XXXX)
python - "$D" <<'PY'
import sys, lzma
d = sys.argv[1]
open(d + "/x.sql", "wb").write(lzma.decompress(open("tests/golden/analytics-hits.sql.xz", "rb").read()))
open(d + "/x.desc.sql", "wb").write(open("tests/golden/analytics-hits.desc.sql", "rb").read())
PY
export PATH=$PWD/oracle/_ref/nocomp:$PATH ZDW_HOST_TIMING=1
cd $D
for i in 1 2; do
T0=$(date +%s.%N); /root/repo/zdw_b200/bin/convertDWfile -q x.sql 2>&1 | tail; python3 -c "import time,sys; print('encode wall %.3f s' % (time.time() - float(sys.argv[1])))" $T0
T0=$(date +%s.%N); /root/repo/zdw_b200/bin/unconvertDWfile -q -d /tmp x.zdw.gz 2>&1 | tail; python3 -c "import time,sys; print('decode wall %.3f s' % (time.time() - float(sys.argv[1])))" $T0
done
T0=$(date +%s.%N); python3 -c "
import ctypes, time
t=time.time(); L=ctypes.CDLL('/root/repo/zdw_b200/libzdw_b200.so'); print('dlopen %.3f'%(time.time()-t))
t=time.time(); h=ctypes.c_void_p(); L.zdwb_ctx_create(0,0,ctypes.byref(h)); print('ctx_create %.3f'%(time.time()-t))
t=time.time(); h2=ctypes.c_void_p(); L.zdwb_ctx_create(0,0,ctypes.byref(h2)); print('2nd ctx_create %.3f'%(time.time()-t))
"
nvidia-smi -q | grep -i "persistence" | head -2
cd /; rm -rf $D
