#!/bin/bash
# N-GPU run (gpurun --gpus N): mix-down tests, then the bench lines as the driver launches them -> gpurun_out/$TAG/
N=${1:-2}; TAG=${2:-r01_n$N}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_mixdown.py -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_mixdown.txt
run() { # name port extra-args...
  local name=$1 port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - $OUT/$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], "|", d["config"].get("mixdown"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -3 $OUT/$name.err
}
KB_MIXDOWN=peer run bench_n$N 29511 --no-extras
KB_MIXDOWN=nccl run bench_n${N}_nccl 29512 --no-extras
KB_MIXDOWN=peer run bench_c5_n$N 29513 --workload c5
KB_MIXDOWN=nccl run bench_c5_n${N}_nccl 29514 --workload c5
