#!/bin/bash
OUT=gpurun_out/${1:-trace}
mkdir -p $OUT
export PYTHONUNBUFFERED=1
KB_RV_TRACE=1 timeout 120 python - > $OUT/trace.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, klang_b200 as kb
fx = kb.FxBank(kb.FX_REVERB, 64, 48000.0, 4096)
io = torch.rand(64, 2, 4096, device="cuda") - 0.5
for _ in range(3):
    fx.process_inplace(io)
torch.cuda.synchronize()
PY
tail -100 $OUT/trace.txt
timeout 120 python tools/fx_probe.py reverb 4096
