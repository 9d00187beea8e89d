#!/bin/bash
# Round-2 capture: GPU tests, smoke, bench (both arms), probes, per-role traces, launch list, ncu --set full of the dominant kernels,
# compute-sanitizer -> gpurun_out/$TAG/
TAG=${1:-r02_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
echo "== tests"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
tail -c 300 $OUT/bench.json; echo
echo "== probes"
{
timeout 120 python tools/fx_probe.py reverb 4096
timeout 120 python tools/fx_probe.py reverb 4096 --tol
timeout 120 python tools/fx_probe.py reverb 4096 --allbus
timeout 120 python tools/fx_probe.py reverb 4096 --tol --allbus
KB_RV_SCHEDULE=2 timeout 120 python tools/fx_probe.py reverb 4096
timeout 120 python tools/fx_probe.py reverb 16384
timeout 120 python tools/fx_probe.py reverb 16384 --tol
timeout 120 python tools/fx_probe.py pingpong 4096
KB_PP_SCHEDULE=2 timeout 120 python tools/fx_probe.py pingpong 4096
timeout 120 python tools/fx_probe.py dreverb 4096
timeout 120 python tools/fx_probe.py dpingpong 4096
timeout 120 python tools/fx_probe.py dpingpong 65536
timeout 120 python tools/c2_probe.py sub
timeout 120 python tools/c2_probe.py tb
timeout 120 python tools/c2_probe.py ssaw
} 2>&1 | grep -v "^$" | tee $OUT/probes.txt
echo "== traces"
KB_RV3_TRACE=$OUT/trace_exact.txt timeout 100 python tools/fx_probe.py reverb 4096 --allbus > /dev/null; python tools/rv3_trace.py $OUT/trace_exact.txt > $OUT/trace_exact_summary.txt
KB_RV3_TRACE=$OUT/trace_tol.txt timeout 100 python tools/fx_probe.py reverb 4096 --tol --allbus > /dev/null; python tools/rv3_trace.py $OUT/trace_tol.txt > $OUT/trace_tol_summary.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_launch_run.log 2>&1
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_sub_mbar -s 3 -c 1 -o $OUT/prof_sub -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > $OUT/ncu_sub.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_reverb3 -s 3 -c 1 -o $OUT/prof_reverb3 -f python tools/fx_probe.py reverb 4096 --allbus > $OUT/ncu_rv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_reverb3 -s 6 -c 1 -o $OUT/prof_reverb3_tol -f python tools/fx_probe.py reverb 4096 --tol --allbus > $OUT/ncu_rvt.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_pingpong3 -s 12 -c 1 -o $OUT/prof_pingpong3 -f python tools/fx_probe.py pingpong 4096 > $OUT/ncu_pp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_mix_fused -s 3 -c 1 -o $OUT/prof_mix_fused -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > $OUT/ncu_mix.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kb_gain_kernel -s 2 -c 1 -o $OUT/prof_gain -f python tools/stream_probe.py > $OUT/ncu_gain.log 2>&1
echo "== stream probe, C2 variants, C2 trace"
timeout 300 python tools/stream_probe.py 2>&1 | tee $OUT/stream_probe.txt
timeout 300 python tools/stream_probe.py 4 2>&1 | tee $OUT/stream_probe_x4.txt
timeout 600 bash tools/bench_ab.sh 2>&1 | tee $OUT/c2_step_ab.txt
KB_C2_TRACE=$OUT/c2_trace.txt timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > /dev/null 2>&1; python tools/c2_trace.py $OUT/c2_trace.txt --raw | tee $OUT/c2_trace_summary.txt | head -8
timeout 120 tools/micro/env_floor > $OUT/env_floor.txt 2>&1
timeout 600 python tools/c2_ab.py 2>&1 | tee $OUT/c2_ab.txt
echo "== sanitizer"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $OUT/compute_sanitizer_memcheck.log 2>&1; tail -3 $OUT/compute_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $OUT/compute_sanitizer_racecheck.log 2>&1; tail -3 $OUT/compute_sanitizer_racecheck.log
ls -la $OUT
