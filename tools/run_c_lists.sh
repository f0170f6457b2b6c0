cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bounds.py -m gpu -x -q -k "stab or long_ranges or capacity or bounds" 2>&1 | tail -3
run() { python bench.py --workload C --no-also 2>/dev/null | grep '^{' > gpurun_out/$1.json; }
run c_lists
BCU_LONG_LISTS=0 run c_nolists
python - <<'PY'
import json,glob
for f in ("gpurun_out/c_lists.json","gpurun_out/c_nolists.json"):
    try:
        d=json.load(open(f)); print(f, d["ms_per_step"], d["roofline"]["frac"], d["build"]["ms"], d["config"]["index"]["device_bytes"])
    except Exception as e: print(f, "ERR", e)
PY
