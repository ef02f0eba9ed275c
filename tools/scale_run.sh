# usage: scale_run.sh N  -- bit-for-bit check + ne120 bench on N GPUs of this box
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_check.py 30 5 11 1 2>&1 | grep -E "mgpu_check|Error|error" | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 16 --warmup 3 > gpurun_out/scale_ne120_n$N.json 2> gpurun_out/scale_ne120_n$N.err
tail -2 gpurun_out/scale_ne120_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/scale_ne120_n$N.json')); print('N=$N', round(d['value'],1), 'tracer-steps/s  ms/tracer-step', round(d['ms_per_tracer_step'],2), 'step frac', round(d['step_hbm']['frac'],3), 'e2e', round(d['e2e']['value'],1)); print(d['timers_ms'])"
