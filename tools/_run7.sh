mkdir -p gpurun_out/s3h
run() { echo -n "$1: "; env $1 timeout 120 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e ms %.4f e2e %.3e launches %s kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms']))"; }
{
run "KB_PDL=0"
run "KB_PDL=1"
run "KB_PDL=0"
run "KB_PDL=1"
} 2>&1 | tee gpurun_out/s3h/pdl_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -x -q -m gpu 2>&1 | tail -3
