# usage: quickbench.sh <ne> <steps>
python bench.py --ne $1 --steps $2 --warmup 2 --no-cpu --no-e2e > gpurun_out/qb.json 2> gpurun_out/qb.err; tail -3 gpurun_out/qb.err
python -c "
import json; d=json.load(open('gpurun_out/qb.json')); print('ne$1', round(d['value'],1), 'ms/tracer-step', round(d['ms_per_tracer_step'],2), 'step frac', round(d['step_hbm']['frac'],3), 'stage avg ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],3)); print(d['timers_ms'])"
