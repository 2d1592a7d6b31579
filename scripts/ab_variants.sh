#!/bin/bash
# A/B of tuning builds on one box: bench.py device time per variant (M2S_LIB selects the library).
# usage: scripts/ab_variants.sh <layout> <variant.so>...   ("default" = the in-tree build)
layout=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then lib=""; else lib="$PWD/$v"; fi
  for rep in 1 2; do
    M2S_LIB=$lib timeout 200 python bench.py --layout $layout --steps 40 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v', '$layout', round(d['ms_per_step'] * 1e3, 2), 'us', round(d['value'], 0), 'Mg/s')"
  done
done
