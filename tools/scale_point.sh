# developer script: one weak-scaling point (run under gpurun --gpus N): bash tools/scale_point.sh N
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --no-extras --no-parity --no-cpu-baseline > gpurun_out/sc_n${N}.json 2> gpurun_out/sc_n${N}.err
python - gpurun_out/sc_n${N}.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n", d['n_gpus'], "value", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "inflight", d['config']['scenes_in_flight'], "cores/rank", d['config'].get('host_cores_per_rank'))
PY
