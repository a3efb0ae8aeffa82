#!/bin/bash
# Usage: tools/dev_variants.sh libssnode libssnode_x ...  -> gpurun_out/variants.log (sweep timing + short bench per library variant)
out=gpurun_out/variants.log; : > $out
for lib in "$@"; do
  echo "== $lib" >> $out
  SSN_LIBNAME=$lib timeout 60 python tools/dev_time_iter.py 2>&1 | tail -1 >> $out
  SSN_LIBNAME=$lib timeout 150 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print('bench: %.0f solves/s  %.2f ms/step  e2e %.0f  frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
" >> $out
done
