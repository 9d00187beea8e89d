#!/bin/bash
TAG=${1:-r1f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== micro"; timeout 60 tools/micro/serial_floor 2>&1 | tee $OUT/micro.txt
echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== probes"
{
for sch in 2; do echo "reverb schedule $sch"; KB_RV_SCHEDULE=$sch timeout 120 python tools/fx_probe.py reverb 4096; done
KB_RV_SCHEDULE=2 timeout 120 python tools/fx_probe.py reverb 16384
} 2>&1 | grep -v "^$" | tee $OUT/probes.txt
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_reverb_pipe -s 3 -c 1 -o $OUT/prof_reverb_pipe -f python tools/fx_probe.py reverb 4096 > $OUT/ncu_rv.log 2>&1
ls -la $OUT
