B="python bench.py --steps 8 --no-extras --no-parity --no-cpu-baseline"
ex() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d['value'],1), "e2e", round(d['e2e']['value'],1), "inflight", d['config']['scenes_in_flight'])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
$B > gpurun_out/g_def.json 2>/dev/null; ex gpurun_out/g_def.json
PCAB_SWITCH_INTERVAL=0.0002 $B > gpurun_out/g_sw.json 2>/dev/null; ex gpurun_out/g_sw.json
PCAB_SWITCH_INTERVAL=0.00002 $B > gpurun_out/g_sw2.json 2>/dev/null; ex gpurun_out/g_sw2.json
PCAB_SWITCH_INTERVAL=0.0002 $B --in-flight 5 > gpurun_out/g_sw5.json 2>/dev/null; ex gpurun_out/g_sw5.json
