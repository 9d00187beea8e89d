#!/bin/bash
# Round capture: GPU tests, bench (both arms), launch list, ncu --set full of the dominant kernels -> gpurun_out/$TAG/
TAG=${1:-r01_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
tail -c 400 $OUT/bench.json; echo
echo "== probes"
{
timeout 120 python tools/fx_probe.py reverb 4096
timeout 120 python tools/fx_probe.py reverb 16384
KB_RV_SCHEDULE=1 timeout 120 python tools/fx_probe.py reverb 4096
timeout 120 python tools/fx_probe.py pingpong 4096
timeout 120 python tools/fx_probe.py dreverb 4096
timeout 120 python tools/fx_probe.py dpingpong 65536
timeout 120 python tools/c2_probe.py sub
KB_TILE_LAYOUT=0 timeout 120 python tools/c2_probe.py sub
timeout 120 python tools/c2_probe.py tb
timeout 120 python tools/c2_probe.py ssaw
timeout 60 tools/micro/serial_floor
} 2>&1 | grep -v "^$" | tee $OUT/probes.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_launch_run.log 2>&1
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_sub_tiled -s 3 -c 1 -o $OUT/prof_sub -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > $OUT/ncu_sub.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_reverb_pipe -s 3 -c 1 -o $OUT/prof_reverb_pipe -f python tools/fx_probe.py reverb 4096 > $OUT/ncu_rv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_dpingpong_stream -s 3 -c 1 -o $OUT/prof_dpp -f python tools/fx_probe.py dpingpong 65536 > $OUT/ncu_dpp.log 2>&1
ls -la $OUT
