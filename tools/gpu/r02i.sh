nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/r02i_bench2.err | tail -1 > gpurun_out/r02i_bench_2gpu.json
python -c "
import json; d=json.loads(open('gpurun_out/r02i_bench_2gpu.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency_mode'), d['init'])"
tail -5 gpurun_out/r02i_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference 2>/dev/null | tail -1 | cut -c1-300
