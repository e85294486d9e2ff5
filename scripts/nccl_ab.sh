N=$1
for ch in default 1 2; do
  if [ "$ch" != "default" ]; then export NCCL_MAX_NCHANNELS=$ch NCCL_MIN_NCHANNELS=1; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((RANDOM%9)) bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=$N channels=$ch value %.3e ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
