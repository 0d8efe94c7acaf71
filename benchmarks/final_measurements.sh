set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r1z3_bench.json 2> gpurun_out/r1z3_bench.err; tail -c 300 gpurun_out/r1z3_bench.json
python benchmarks/ntt_micro.py --logs 13,16,18,20,21,22,23 --lde > gpurun_out/r1z3_ntt_micro_warm.jsonl 2>&1
python benchmarks/ntt_micro.py --logs 16,18,20,22 --batch 16 --lde > gpurun_out/r1z3_ntt_micro_batch16.jsonl 2>&1
python benchmarks/cfg2_univariate.py > gpurun_out/r1z3_cfg2_univariate.jsonl 2>&1; tail -4 gpurun_out/r1z3_cfg2_univariate.jsonl | cut -c1-200
python benchmarks/sharded_sweep.py --sizes 22 --cfg4-log-n 22 --check > gpurun_out/r1z3_sweep_g1.jsonl 2>&1; tail -3 gpurun_out/r1z3_sweep_g1.jsonl | cut -c1-300
python bench.py --cols 16 --no-cpu-baseline > gpurun_out/r1z3_bench_cols16.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r1z3_bench_cols16.json')); print('cols16', d['ms_per_step'], d['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1z3_launches_bench_one_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
