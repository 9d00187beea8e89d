#!/bin/bash
TAG=${1:-r1n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== probes"
{
for cfg in "0 544 7" "2 768 8"; do
  set -- $cfg
  echo "layout $1 nt $2 g $3"
  KB_TILE_LAYOUT=$1 KB_TILE_NT=$2 KB_TILE_G=$3 timeout 120 python tools/c2_probe.py sub
  KB_TILE_LAYOUT=$1 KB_TILE_NT=$2 KB_TILE_G=$3 timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  bench: value %.3e e2e %.3e ms_per_step %.4f kernel_ms %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms']))"
done
} 2>&1 | grep -v "^$" | tee $OUT/probes.txt
echo "== ncu full"
KB_TILE_LAYOUT=2 KB_TILE_G=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_sub_tiled -s 3 -c 1 -o $OUT/prof_sub_l2 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > $OUT/ncu_sub.log 2>&1
ls -la $OUT | tail -4
