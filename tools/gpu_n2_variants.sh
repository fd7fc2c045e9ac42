# 2-GPU attribution of the exchange cost: fit only / peer stores without signalling / + signal / + lagged wait
mkdir -p gpurun_out
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-extra --no-e2e --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', 'step %.4f ms  fit-only(eager) %.4f  K3 %.4f  peer_check %s'%(d['ms_per_step'], 2*d['config']['pixels_per_gpu']/d['fit_only_px_s']*1e3, d['roofline']['kernel_ms'], d['peer_check'] and d['peer_check']['bit_identical']))"; }
run MT_PEER_SYNC=none
run MT_PEER_SYNC=signal
run MT_PEER_SYNC=full
run MT_PEER_SYNC=full MT_X=--gather
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-extra --no-e2e --no-cpu-baseline --steps 40 --no-graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('eager (no graph) step %.4f ms'%d['ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --no-extra --no-e2e --no-cpu-baseline --steps 40 --nccl-gather 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('NCCL all-gather step %.4f ms'%d['ms_per_step'])"
python bench.py --no-extra --no-e2e --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1 GPU step %.4f ms K3 %.4f'%(d['ms_per_step'], d['roofline']['kernel_ms']))"
