set -x
timeout 120 python profiles/kernel_ab.py --steps 20 2>&1 | tail -1 | tee gpurun_out/r2au_ilv_ab.txt
EMC_LIB=$PWD/variants/libemc_noilv.so timeout 120 python profiles/kernel_ab.py --steps 20 2>&1 | tail -1 | tee -a gpurun_out/r2au_ilv_ab.txt
timeout 120 python profiles/kernel_ab.py --steps 20 --no-fused 2>&1 | tail -1 | tee -a gpurun_out/r2au_ilv_ab.txt
EMC_LIB=$PWD/variants/libemc_noilv.so timeout 120 python profiles/kernel_ab.py --steps 20 --no-fused 2>&1 | tail -1 | tee -a gpurun_out/r2au_ilv_ab.txt
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2au_pytest.txt
