# chained 2c -> next 2a launches (SBB_CHAIN, default 1): bit-identity test, then same-box ABAB per layer
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "chained or image_spanning" 2>&1 | tail -15 | tee gpurun_out/r02ag_tests.txt
export F='page 2800|res3b_branch2c|res3c_branch2a|res3c_branch2c|res3d_branch2a|res4b_branch2c|res4c_branch2a|res4c_branch2c|res5b_branch2c|res5c_branch2a|sum of'
VAR=SBB_CHAIN VALS="0 1 0 1" timeout 600 bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02ag_chain_abab.txt
