#!/bin/bash
# usage: bench3.sh tag [env...]
tag=$1; shift
for w in c2 c3 c1; do env "$@" python bench.py --workload $w --steps 50 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/${tag}_$w.json 2>gpurun_out/${tag}_$w.err; python -c "
import json,sys
j=[json.loads(l) for l in open('gpurun_out/${tag}_$w.json') if l.startswith('{')][0]
print('$tag','$w', round(j['ms_per_step'],4), round(j['value'],3), j.get('pass_ms'), round(j['e2e']['value'],3))
"; done
