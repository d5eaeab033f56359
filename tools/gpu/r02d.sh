timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "pair" 2>&1 | tail -15
VAR=SBB_PAIR VALS="0 1 0 1" F='page 2800|res3b_branch2b|res4b_branch2b|res4b_branch2c|res5b_branch2b|res5b_branch2a|dec_v4|dec1|dec2|dec3|sum of' timeout 400 bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02d_pair_abab.txt
