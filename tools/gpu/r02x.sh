timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -k "pair or every_layer or tile_logits or golden_tile or simt" 2>&1 | tail -4
VAR=SBB_PAIR_RESB VALS="0 1 0 1" F='page 2800|conv1 |res2a_branch2b|res2b_branch2a|res2b_branch2b|sum of' timeout 400 bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02x_resident_b_abab.txt
