#!/bin/bash
# one GPU call: tests (with the PARITY lines), the bench line, the launch list of the bench command, per-model timings, the ncu session
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -s 2>&1 > gpurun_out/pytest_s.log; grep PARITY gpurun_out/pytest_s.log > gpurun_out/r02_parity_gpu.txt; tail -3 gpurun_out/pytest_s.log > gpurun_out/r02_pytest_gpu.log; rm gpurun_out/pytest_s.log
cat gpurun_out/r02_pytest_gpu.log
timeout 500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1.json').read()); print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); print(d['e2e']); print({k:(round(v['ms_per_step'],2),round(v['frac'],3)) for k,v in d['roofline']['kernels'].items()}); print(d['configs'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_command.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 300 python tools/time_models.py > gpurun_out/r02_all_models.log 2>&1; cat gpurun_out/r02_all_models.log
bash tools/ncu_session.sh > /dev/null 2>&1; cut -c1-200 gpurun_out/r02_kernels_ncu.txt
