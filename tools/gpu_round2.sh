#!/bin/bash
TAG=${1:-r1c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== probes"
{
for cfg in "0 512 8" "1 512 8" "1 768 8" "1 768 7"; do
  set -- $cfg
  echo "layout $1 nt $2 g $3"; KB_TILE_LAYOUT=$1 KB_TILE_NT=$2 KB_TILE_G=$3 timeout 120 python tools/c2_probe.py sub
done
for sch in 1 2; do echo "reverb schedule $sch"; KB_RV_SCHEDULE=$sch timeout 120 python tools/fx_probe.py reverb 4096; done
KB_RV_SCHEDULE=2 timeout 120 python tools/fx_probe.py reverb 16384
} 2>&1 | grep -v "^$" | tee $OUT/probes.txt
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_reverb_pipe -s 3 -c 1 -o $OUT/prof_reverb_pipe -f python tools/fx_probe.py reverb 4096 > $OUT/ncu_rv.log 2>&1
KB_TILE_LAYOUT=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_sub_tiled -s 3 -c 1 -o $OUT/prof_sub_l0 -f python tools/c2_probe.py sub > $OUT/ncu_sub0.log 2>&1
KB_TILE_LAYOUT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_sub_tiled -s 3 -c 1 -o $OUT/prof_sub_l1 -f python tools/c2_probe.py sub > $OUT/ncu_sub1.log 2>&1
ls -la $OUT
