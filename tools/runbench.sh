python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for ne in 30 120; do
  if [ $ne = 30 ]; then ST=4; else ST=16; fi
  python bench.py --ne $ne --steps $ST --warmup 3 --no-cpu > gpurun_out/bench_ne$ne.json 2> gpurun_out/bench_ne$ne.err; tail -3 gpurun_out/bench_ne$ne.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ne$ne.json')); print('ne$ne', round(d['value'],1), 'ms/tracer-step', round(d['ms_per_tracer_step'],2), 'step frac', round(d['step_hbm']['frac'],3), 'stage avg ms', round(d['roofline']['avg_launch_ms'],3), 'frac', round(d['roofline']['frac'],3)); print(d['timers_ms']); print('e2e', d['e2e']['value'] if d['e2e'] else None)"
done
