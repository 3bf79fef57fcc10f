#!/bin/bash
# round 2, GPU session N (1 GPU): trainer e2e + loader + train tests, ncu of the two step kernels at the reference
# batch (2048 + 2048), loader timing probe, headline check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer_e2e.py tests/test_gpu_loader.py tests/test_gpu_train.py tests/test_gpu_sampler.py -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -n 4 gpurun_out/r2n_pytest.log
B="python bench.py --steps 10 --warmup 5 --no-extras --no-cpu-baseline"
$B | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("cfg2 ms/step", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam", d["roofline"]["adam_ms"])'
$B --workload cfg2_transe_ml1m_b2048 > gpurun_out/r2n_plain_b2048.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"train_fwd|adam_apply" -s 20 -c 2 -f -o /tmp/r2n_prof_b2048 $B --workload cfg2_transe_ml1m_b2048 > gpurun_out/r2n_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/r2n_prof_b2048.ncu-rep gpurun_out/r2_train_cfg2_b2048.txt --top 30 | tail -1
cp /tmp/r2n_prof_b2048.ncu-rep gpurun_out/ 2>/dev/null
timeout 300 python scripts/loader_probe.py > gpurun_out/r2n_loader_probe.log 2>&1; cat gpurun_out/r2n_loader_probe.log
du -sh gpurun_out
