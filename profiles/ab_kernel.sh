#!/bin/bash
# gpurun -- 'bash profiles/ab_kernel.sh tag1 tag2 ...'   (variants/libemc_<tag>.so; "default" = the in-tree build)
# per variant: kernel time + state checksum (kernel_ab.py), then executed warp-instructions per launch (ncu)
for v in "$@"; do
  if [ "$v" = default ]; then unset EMC_LIB; else export EMC_LIB=$PWD/variants/libemc_$v.so; fi
  python profiles/kernel_ab.py --tag $v $AB_ARGS 2>/dev/null | tail -1
  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum -k regex:flight_scatter -c 2 \
      --csv python profiles/kernel_ab.py --steps 2 --warmup 0 2>/dev/null | grep -E "inst_executed|time_duration" | awk -F'","' '{print "   ncu", $(NF-2), $NF}' | tr -d '"' | sort -u | head -6
done
