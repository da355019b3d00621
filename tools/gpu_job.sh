#!/bin/bash
mkdir -p gpurun_out
bash tools/bench_all.sh C2 C4 C5-512
BENCH_ENV="AW_L2_PERSIST=0" bash tools/bench_all.sh C2 C4 C5-512
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persistingL2CacheMaxSize", getattr(p, "persisting_l2_cache_max_size", None))
PY
