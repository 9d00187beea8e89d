# A/B of the C2 step variants on the real bench (measurement aid): bash tools/bench_ab.sh
run() { echo -n "$1: "; env $1 timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e ms %.4f e2e %.3e launches %s kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms']))"; }
run KB_TILE_LAYOUT=2
run KB_TILE_LAYOUT=3
run KB_TILE_LAYOUT=4
run "KB_TILE_LAYOUT=4 KB_PDL=0"
run "KB_TILE_LAYOUT=4 KB_STAGE_COPY=0"
run "KB_TILE_LAYOUT=4 KB_ZERO_COPY_OUT=0"
run "KB_TILE_LAYOUT=4 KB_TILE_G=8"
run "KB_TILE_LAYOUT=4 KB_C2_VARIANT=64"
run "KB_TILE_LAYOUT=4 KB_C2_VARIANT=128"
run "KB_TILE_LAYOUT=3 KB_SCATTER_FUSED=0"
run "KB_TILE_LAYOUT=3 KB_TILE_ASP0=0"
run "KB_TILE_LAYOUT=3 KB_TILE_G=8"
run "KB_TILE_LAYOUT=3 KB_MIX_FUSED=0"
