#!/bin/bash
# A/B of tuning variants of the flight kernel on ONE box (same clocks, same command):
#   EMC_NVCC_EXTRA="-DEMC_SPEC_SELECT=0" EMC_LIB_OUT=$PWD/variants/libemc_nospec.so \
#       python -m monte_carlo_carrier_transport_b200.build --force        # one .so per variant (variants/ is git-ignored)
#   gpurun -- 'bash profiles/ab_variants.sh nospec default'
# Knobs: EMC_FAST_MATH_PATHS, EMC_FAST_SAMPLERS, EMC_SPEC_SELECT, EMC_BRANCHFREE_INDEX,
# EMC_LOG_CONSTANT_BANK, EMC_STAGE_REFILL, EMC_PREFETCH_BATCHES, EMC_FIELD_AHEAD, EMC_OBS_SMEM,
# EMC_FLIGHT_THREADS / EMC_FLIGHT_MIN_BLOCKS / EMC_FLIGHT_MAXNREG (compile time); EMC_ZONLY=0,
# EMC_FIELD_AHEAD=0 and `bench.py --obs-phase after` (run time).  Logs: profiles/r1*_ab_*.txt.
for v in "$@"; do
  if [ "$v" = default ]; then unset EMC_LIB; else export EMC_LIB=$PWD/variants/libemc_$v.so; fi
  for rep in 1 2; do
    python bench.py --steps 30 --warmup 5 --no-coupled 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', '%.4g steps/s  kernel %.4f ms  step %.4f ms  frac %.3f  faults %d' % (d['value'], d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'], d['faulted_particles']))
"
  done
done
