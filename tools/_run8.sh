mkdir -p gpurun_out/s3i
run() { echo -n "$1: "; env $1 timeout 120 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e ms %.4f e2e %.3e launches %s kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms']))"; }
V() { echo $(( (($1+1)<<8) | (($2+1)<<16) )); }
{
run "KB_C2_VARIANT=0"
run "KB_C2_VARIANT=$(V 20 19)"
run "KB_C2_VARIANT=$(V 20 23)"
run "KB_C2_VARIANT=$(V 20 3)"
run "KB_C2_VARIANT=$(V 20 1)"
run "KB_C2_VARIANT=0 KB_TILE_G=8"
} 2>&1 | tee gpurun_out/s3i/roles_ab.txt
