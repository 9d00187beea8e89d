mkdir -p gpurun_out/s3f
run() { echo -n "$1: "; env $1 timeout 120 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e ms %.4f e2e %.3e launches %s kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms']))"; }
V() { echo $(( (($1+1)<<8) | (($2+1)<<16) )); }
timeout 300 python tools/c2_ab.py 2>&1 | tail -3 | tee gpurun_out/s3f/c2_ab.txt
{
run "KB_C2_VARIANT=0"
run "KB_C2_VARIANT=$(V 20 1)"
run "KB_TILE_LAYOUT=3"
} 2>&1 | tee gpurun_out/s3f/bench_ab.txt
KB_C2_TRACE=gpurun_out/s3f/tr.txt timeout 120 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > /dev/null 2>&1; python tools/c2_trace.py gpurun_out/s3f/tr.txt --raw > gpurun_out/s3f/tr_summary.txt; head -26 gpurun_out/s3f/tr_summary.txt
KB_C2_VARIANT=$(V 20 1) KB_C2_TRACE=gpurun_out/s3f/tr_a1.txt timeout 120 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > /dev/null 2>&1; python tools/c2_trace.py gpurun_out/s3f/tr_a1.txt | tee gpurun_out/s3f/tr_a1_summary.txt
