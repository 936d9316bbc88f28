#!/bin/bash
# usage (GPU box): tools/ab1.sh name1 name2 ...  — c3 and c2 only (30 steps), once per variant
mkdir -p gpurun_out
for v in "$@"; do for w in c3 c2; do UVT_NO_REBUILD=1 UVT_LIB_PATH=variants/libuvt_$v.so python bench.py --workload $w --steps 40 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ab1_${v}_$w.json 2>/dev/null; python -c "
import json
j=[json.loads(l) for l in open('gpurun_out/ab1_${v}_$w.json') if l.startswith('{')][0]
print('ab1','$v','$w', round(j['ms_per_step'],4), j.get('pass_ms'))
"; done; done
