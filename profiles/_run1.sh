for v in anchor1 obssmem; do EMC_LIB=$PWD/variants/libemc_$v.so python profiles/kernel_ab.py --no-fused --tag $v 2>&1 | tail -1; done
EMC_BULK_REFILL=0 EMC_LIB=$PWD/variants/libemc_thr_640o.so python profiles/kernel_ab.py --no-fused --tag thr_640o_nobulk 2>&1 | tail -1
EMC_BULK_REFILL=0 EMC_LIB=$PWD/variants/libemc_thr_640o.so ncu --set full --clock-control none --import-source on -k regex:flight_scatter -s 3 -c 1 -o gpurun_out/flight_r2t_640o -f python profiles/kernel_ab.py --no-fused --steps 2 --warmup 2 > gpurun_out/ncu_full_r2t_640o.log 2>&1
