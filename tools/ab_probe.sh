#!/bin/bash
# same-box A/B of builds of the library on the production null (C3): SAFE_B200_LIB selects the build
run() {
  env "$@" python bench.py --no-ceiling --no-cpu-baseline --no-safe-api --no-parity --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],1), {k: round(v,1) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
}
for v in "$@"; do run SAFE_B200_LIB=$PWD/safepy_b200/libsafe_b200_$v.so; done
