# usage: scale_run.sh N [quick]  -- bit-for-bit check + benches on N GPUs of this box (ne120 DCMIP 1-1, ne120 DCMIP 1-2, ne30)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29531 tests/mgpu_check.py 30 5 11 1 2>&1 | grep -E "mgpu_check|differs|Error|error" | tail -3
run() {  # tag, bench args...
  tag=$1; shift
  $TR --master-port 29532 bench.py --gpus $N "$@" > gpurun_out/scale_${tag}_n$N.json 2> gpurun_out/scale_${tag}_n$N.err || tail -3 gpurun_out/scale_${tag}_n$N.err
  python -c "
import json; d=json.load(open('gpurun_out/scale_${tag}_n$N.json')); print('$tag N=$N', round(d['value'],1), 'tracer-steps/s  ms/tracer-step', round(d['ms_per_tracer_step'],2), 'step frac', round(d['step_hbm']['frac'],3), 'e2e', round(d['e2e']['value'],1) if d['e2e'] else None, 'hash', d['field_hash_all'], d['timers_ms'])"
}
run ne120_t11 --steps 16 --warmup 3 --no-cpu
if [ "$2" != quick ]; then
  run ne120_t12 --test 12 --steps 8 --warmup 2 --no-cpu --no-e2e
  run ne30_t11 --ne 30 --steps 16 --warmup 3 --no-cpu --no-e2e
fi
