for v in default ilvmesh default ilvmesh; do
  if [ "$v" = default ]; then unset EMC_LIB; else export EMC_LIB=$PWD/variants/libemc_$v.so; fi
  timeout 120 python profiles/kernel_ab.py --steps 20 --no-fused --mesh --tag $v 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['tag'], d['configs1']['kernel_ms_mean'], d['coupled_mesh'])" | tee -a gpurun_out/r2aw_ilv_mesh_ab.txt
done
