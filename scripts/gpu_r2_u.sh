#!/bin/bash
# round 2, GPU session U (1 GPU): whole GPU suite (TransH fix, int32 id staging), bench N=1, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
grep -E "^E  |passed|failed|FAILED|rc=" gpurun_out/r2u_pytest.log | head -n 30 | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam", d["roofline"]["adam_ms"], "e2e", d["e2e"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_bench_ref.json").read().strip().splitlines()[-1])
print("ref value", d["value"], "ms", d["ms_per_step"], d["cpu_baseline"]["cores"], json.dumps(d.get("extras"))[:600])
PY
