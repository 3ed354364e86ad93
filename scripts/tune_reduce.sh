timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step']*1000, d['roofline']['kernel_timed_alone_us'], {k:round(v['us_per_step'],1) for k,v in d['roofline']['kernels_us_per_step'].items()})"; }
run probe8 A=1
run probe6 PERSIA_B200_LIB=/root/repo/persia_b200/libpersia_b200_pb6.so
run probe5 PERSIA_B200_LIB=/root/repo/persia_b200/libpersia_b200_pb5.so
