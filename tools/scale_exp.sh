# developer script: 8-GPU weak-scaling probe with different numbers of scenes in flight (run under gpurun --gpus 8)
N=${1:-8}
B="bench.py --gpus $N --steps 6 --no-extras --no-parity --no-cpu-baseline"
ex() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n", d['n_gpus'], "value", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "inflight", d['config']['scenes_in_flight'], d['config'].get('host_sync'), "cores/rank", d['config'].get('host_cores_per_rank'), "serial", round(d['config']['serial_ms_per_forward'],2))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
nproc
python bench.py --steps 6 --no-extras --no-parity --no-cpu-baseline > gpurun_out/sc_n1.json 2>/dev/null; ex gpurun_out/sc_n1.json
for IF in 0 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 $B --in-flight $IF > gpurun_out/sc_n${N}_if$IF.json 2> gpurun_out/sc_n${N}_if$IF.err; ex gpurun_out/sc_n${N}_if$IF.json
done
