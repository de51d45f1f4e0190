# The commands behind profiles/r2z_* (two gpurun calls; summarize_ncu.py runs on the build host in between, so that
# the bench line's static roofline.traffic comes from the capture of the same build):
#   gpurun -- 'bash profiles/_final.sh ncu'   ; python profiles/summarize_ncu.py r2z --rep gpurun_out/flight_r2z.ncu-rep --launches gpurun_out/r2z_launches.csv
#   gpurun -- 'bash profiles/_final.sh'       ; cp gpurun_out/r2z_{bench,bench_reference}.json gpurun_out/r2z_pytest_gpu.log profiles/
set -x
if [ "$1" = ncu ]; then
  python bench.py --steps 5 --warmup 3 --no-e2e --no-coupled > gpurun_out/r2z_noe2e.json 2> gpurun_out/r2z_noe2e.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-coupled > gpurun_out/r2z_ncu_launch.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:flight_scatter -s 4 -c 1 -o gpurun_out/flight_r2z -f python profiles/kernel_ab.py --no-fused --steps 2 --warmup 3 > gpurun_out/r2z_ncu_full.log 2>&1
  exit 0
fi
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2z_pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -c 300 gpurun_out/r2z_bench.json; cat gpurun_out/r2z_pytest_gpu.log; tail -3 gpurun_out/r2z_bench.err
