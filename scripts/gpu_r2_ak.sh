#!/bin/bash
# round 2, GPU session AK (1 GPU): touched-row lists for small batches -- train tests, reference-batch timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_loader.py tests/test_gpu_trainer_e2e.py -m gpu -q 2>&1 | tail -n 6
python - <<'PY'
import subprocess, json
out = subprocess.run("python bench.py --steps 20 --warmup 5 --no-cpu-baseline", shell=True, capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1])
print("headline", d["ms_per_step"], d["roofline"]["fwd_ms"], d["roofline"]["adam_ms"])
for k in ("cfg2_transe_ml1m_b2048", "cfg5_transe_alibaba_b2048", "cfg5_transe_alibaba", "cfg3_rotate_yelp"):
    x = d["extras"][k]
    print(k, round(x["ms_per_step"] * 1e3, 1), "fwd", round(x["fwd_ms"] * 1e3, 1), "adam", round(x["adam_ms"] * 1e3, 1), x.get("train_step_call", {}).get("device_ms_per_step"))
print(json.dumps(d["extras"]["cfg2_b2048_device_loader"]))
PY
