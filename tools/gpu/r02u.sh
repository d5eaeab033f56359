timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -k "pair or every_layer or tile_logits or golden_tile" 2>&1 | tail -4
VAR=SBB_PAIR64 VALS="0 1 0 1" F='page 2800|conv1 |res2a_branch2b|res2b_branch2a|res2b_branch2b|sum of' timeout 400 bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02u_pair64_abab.txt
python bench.py --no-cpu-baseline --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['parts'], d['roofline']['parts_how'])"
