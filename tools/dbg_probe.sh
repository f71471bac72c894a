#!/bin/bash
# timing experiments on the production null (C3).  Results with SB_DBG & (1|2|4|8|16) are garbage.
run() {
  env "$@" python bench.py --no-cpu-baseline --no-safe-api --no-parity --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
}
for dbg in 0 8 16 24; do run SB_DBG=$dbg; done
