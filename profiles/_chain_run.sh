for v in default rotate tiles2 pushlast parkf0 default rotate tiles2 pushlast parkf0; do
  if [ "$v" = default ]; then unset EMC_LIB; else export EMC_LIB=$PWD/variants/libemc_$v.so; fi
  timeout 120 python profiles/kernel_ab.py --steps 30 --no-fused --tag $v 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['tag'], d['configs1']['kernel_ms_mean'], d['configs1']['kernel_ms_min'], d['configs1']['state_checksum'])" | tee -a gpurun_out/r2ax_ab.txt
done
