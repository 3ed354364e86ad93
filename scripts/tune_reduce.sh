timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline 2>gpurun_out/tune.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step']*1000, d['roofline']['kernel_timed_alone_us'], d['gpu_launches'], {k:round(v['us_per_step'],1) for k,v in d['roofline']['kernels_us_per_step'].items()})" || tail -5 gpurun_out/tune.err; }
run fused A=1
run separate PB_NO_FUSED_GROUPING=1
