# direct st.global epilogue stores of the CTA-pair kernel (SBB_DIRECT_STORE, default 1) vs TMA stores: parity, then same-box ABAB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02ac_tests.txt
export F='page 2800|conv1 |res2a_branch2b|res2a_branch2c|res2b_branch2a|res2b_branch2c|res3b_branch2a|res3b_branch2b|res3b_branch2c|res4b_branch2a|res4b_branch2b|res4b_branch2c|res5b_branch2a|res5b_branch2b|res5b_branch2c|dec_v4|dec1|dec2|dec3|dec4|dec5|sum of'
VAR=SBB_DIRECT_STORE VALS="0 1 0 1" bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02ac_direct_store_abab.txt
