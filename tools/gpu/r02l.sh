nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/r02l_bench8.err | tail -1 > gpurun_out/r02l_bench_8gpu.json
python -c "
import json; d=json.loads(open('gpurun_out/r02l_bench_8gpu.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency_mode'), d['init'], d['clocks'])"
tail -3 gpurun_out/r02l_bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --config 5 2>>gpurun_out/r02l_bench8.err | tail -1 > gpurun_out/r02l_bench_8gpu_cfg5.json
python -c "
import json; d=json.loads(open('gpurun_out/r02l_bench_8gpu_cfg5.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency_mode'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 10 --warmup 3 --config 3 2>>gpurun_out/r02l_bench8.err | tail -1 > gpurun_out/r02l_bench_8gpu_cfg3.json
python -c "
import json; d=json.loads(open('gpurun_out/r02l_bench_8gpu_cfg3.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
tail -3 gpurun_out/r02l_bench8.err
