set -x
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2/bench_n$N.json 2> gpurun_out/r2/bench_n$N.err
echo rc=$?
tail -1 gpurun_out/r2/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); ex=d.get('extras',{})
print({k:d[k] for k in ('n_gpus','value','ms_per_step','host_buffer_placement')}); print(json.dumps(d['e2e'])[:420])
for k,v in ex.items(): print(k, json.dumps(v)[:600])
"
nvidia-smi topo -m > gpurun_out/r2/topo_n$N.txt 2>&1
