# final measurement set of round 2 (1 GPU): tests, config 2 full line, DDIM-20, configs 1 / 3 / 5
set -x
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
cp gpurun_out/parity_report.json gpurun_out/final_parity_report.json
j() { grep '^{"metric' | tail -1; }
python bench.py --steps 20 --warmup 5 2>/dev/null | j > gpurun_out/final_c2.json
python bench.py --steps 10 --warmup 3 --sampler ddim --ddpm_steps 20 --no_cpu_baseline 2>/dev/null | j > gpurun_out/final_c2_ddim20.json
python bench.py --config 1 --steps 10 --warmup 3 --no_cpu_baseline 2>/dev/null | j > gpurun_out/final_c1.json
python bench.py --config 3 --quick --steps 2 --warmup 1 --no_cpu_baseline 2>/dev/null | j > gpurun_out/final_c3.json
python bench.py --config 5 --quick --steps 2 --warmup 1 --no_cpu_baseline 2>/dev/null | j > gpurun_out/final_c5.json
python profiles/stage_times.py --config 2 2>/dev/null | tail -1 > gpurun_out/final_stage_times_c2.json
for f in gpurun_out/final_c*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['value'], d.get('value_depth1'), (d.get('e2e') or {}).get('value'), d['roofline']['frac'], d['config'].get('sweep'))"; done
