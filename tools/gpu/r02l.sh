nvidia-smi -L | wc -l
for cfg in 2 5 3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$cfg bench.py --gpus 8 --steps 20 --warmup 3 --config $cfg 2>gpurun_out/r02z_bench8_$cfg.err | tail -1 > gpurun_out/r02z_bench_8gpu_cfg$cfg.json
python -c "
import json; d=json.loads(open('gpurun_out/r02z_bench_8gpu_cfg$cfg.json').read()); print($cfg, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency_mode'), d['clocks'])"
tail -2 gpurun_out/r02z_bench8_$cfg.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus 8 --steps 2 --warmup 1 --impl reference 2>/dev/null | tail -1 | cut -c1-250
