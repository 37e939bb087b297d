#!/bin/bash
# prints the frame-kernel / predictor split of bench.py for a few settings passed as "ENV=VAL ..." strings
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps 20 --warmup 3 --no-latency --no-cpu-baseline --no-config3 ${BENCH_ARGS} 2>/dev/null)
  python - "$cfg" <<PY
import json,sys
d=json.loads('''$out''')
r=d["roofline"]
print(f"{sys.argv[1]:40s} step {d['ms_per_step']:.4f} ms  frame {r['kernel_ms_per_launch']:.4f}  predictor {r['predictor_ms_per_step']:.4f}  launches/step {d['gpu_launches']/d['steps']:.1f}  e2e {d['e2e']['value']/1e6:.3f} M")
PY
done
