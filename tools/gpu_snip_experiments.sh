mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log); tail -3 gpurun_out/r2e_pytest.log
B="python bench.py --workload c3 --no-extra --no-e2e --no-cpu-baseline --steps 8"
for cfg in "MT_SNIP_PREFETCH=0" "MT_SNIP_PREFETCH=296" "MT_SNIP_PREFETCH=148" "MT_SNIP_PREFETCH=296 MT_SNIP_FASTMATH=1"; do
  env $cfg $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$cfg', 'step %.3f ms  snip %.4f ms  contract %.4f  check'%(d['ms_per_step'], r['kernel_ms'], r['contraction']['kernel_ms']), d['check_vs_float64_lstsq'].get('nnls_max_rel_err_interior'))"
done
MT_SNIP_FASTMATH=1 timeout 300 python -m pytest tests/test_gpu_snip.py tests/test_gpu_c3.py -m gpu -q -s 2>&1 | grep -E "passed|failed|c3 shape|rel|Error" | tail -8
