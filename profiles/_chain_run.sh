set -x
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "fused or chain or host_stepper or config" 2>&1 | tail -5
timeout 120 python profiles/kernel_ab.py --steps 20 2>&1 | tail -1 | tee gpurun_out/r2av_ilv_fused_ab.txt
cat > /tmp/e2e_ab.py <<'P'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from monte_carlo_carrier_transport_b200 import abi, engine
n = bench.N_FIELDS * bench.PER_FIELD
ctx = engine.Context(device=0)
ctx.set_field_groups(bench.field_points(), bench.PER_FIELD)
ens = engine.Ensemble(ctx, n).init(bench.SEED)
co, val, dtau = ens.coords()
del ens
for chunk_log2, ramp_log2 in ((23, 20), (22, 20), (23, 19), (23, 20)):
    hs = engine.HostStepper(ctx, n, chunk=1 << chunk_log2, ramp_from=1 << ramp_log2)
    hs.coords[:], hs.valley[:], hs.dtau[:] = co, val, dtau
    hs.run(0, 50, 1, field_mode=abi.FIELD_GROUPS)
    torch.cuda.synchronize()
    ms = []
    for r in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); hs.run(50 * (r + 1), 50, 1, field_mode=abi.FIELD_GROUPS); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b) / 50)
    print("chain=%s chunk 2^%d ramp 2^%d: %.4f ms per timestep (mean of 3), min %.4f" % (ctx.describe()["chain_steps"], chunk_log2, ramp_log2, np.mean(ms), min(ms)), flush=True)
    del hs
    torch.cuda.empty_cache()
P
timeout 150 python /tmp/e2e_ab.py 2>&1 | tail -4 | tee gpurun_out/r2av_e2e_chain_ab.txt
EMC_CHAIN_STEPS=0 timeout 150 python /tmp/e2e_ab.py 2>&1 | tail -4 | tee -a gpurun_out/r2av_e2e_chain_ab.txt
