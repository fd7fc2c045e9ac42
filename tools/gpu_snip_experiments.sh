# SNIP: contiguous-run kernel (csrc/mt_snip_runs.cu) against the strided one: parity, then the c3 step with variants
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_snip.py tests/test_gpu_c3.py -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/snipexp_pytest.log
B="python bench.py --workload c3 --no-extra --no-e2e --no-cpu-baseline --steps 8"
for cfg in "MT_SNIP_RUNS=1" "MT_SNIP_RUNS=1 MT_SNIP_FASTMATH=1" "MT_SNIP_RUNS=1 MT_SNIP_DEBUG_NPASS=0" "MT_SNIP_RUNS=1 MT_SNIP_DEBUG_NPASS=4" "MT_SNIP_RUNS=1 MT_SNIP_DEBUG_NPASS=8"; do
  env $cfg timeout 300 $B 2>gpurun_out/snipexp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$cfg', 'step %.3f ms  snip %.4f ms per %d px  check %s'%(d['ms_per_step'], r['kernel_ms'], r['px_per_launch'], d['check_vs_float64_lstsq'].get('nnls_max_rel_err_interior')))" 2>&1 | tail -1
done | tee gpurun_out/snipexp.log
