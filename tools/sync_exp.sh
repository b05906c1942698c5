B="python bench.py --steps 6 --no-extras --no-parity --no-cpu-baseline"
ex() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "inflight", d['config']['scenes_in_flight'], d['config'].get('host_sync'), "cores", d['config'].get('host_cores_per_rank'), "serial", round(d['config']['serial_ms_per_forward'],2))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
nproc
PCAB_HOST_SYNC=default taskset -c 0-3 $B --in-flight 3 > gpurun_out/x_t4_spin3.json 2>/dev/null; ex gpurun_out/x_t4_spin3.json
PCAB_HOST_SYNC=blocking taskset -c 0-3 $B --in-flight 3 > gpurun_out/x_t4_block3.json 2>/dev/null; ex gpurun_out/x_t4_block3.json
PCAB_HOST_SYNC=blocking taskset -c 0-3 $B --in-flight 4 > gpurun_out/x_t4_block4.json 2>/dev/null; ex gpurun_out/x_t4_block4.json
PCAB_HOST_SYNC=yield taskset -c 0-3 $B --in-flight 4 > gpurun_out/x_t4_yield4.json 2>/dev/null; ex gpurun_out/x_t4_yield4.json
PCAB_HOST_SYNC=blocking $B --in-flight 4 > gpurun_out/x_all_block4.json 2>/dev/null; ex gpurun_out/x_all_block4.json
PCAB_HOST_SYNC=default $B --in-flight 4 > gpurun_out/x_all_spin4.json 2>/dev/null; ex gpurun_out/x_all_spin4.json
