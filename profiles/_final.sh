set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2z_pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -c 300 gpurun_out/r2z_bench.json; cat gpurun_out/r2z_pytest_gpu.log; tail -3 gpurun_out/r2z_bench.err
