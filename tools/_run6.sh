mkdir -p gpurun_out/s3g
timeout 120 tools/micro/env_floor > gpurun_out/s3g/env_floor.txt; grep -i "tile\|FADD\|FMUL\|lanes" gpurun_out/s3g/env_floor.txt
for g in sub fm senv ssaw tb; do timeout 120 python tools/c2_probe.py $g; done 2>&1 | tee gpurun_out/s3g/probes.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s3g/pytest_gpu.txt
