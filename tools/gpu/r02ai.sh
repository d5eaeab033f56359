timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "chained" 2>&1 | tail -3
bash tools/gpu/r02ah.sh 2>&1 | cut -c1-250 | grep -E "==|\+" | head -8
export F='page 2800|\+res|sum of'
VAR=SBB_CHAIN VALS="0 1 0 1" timeout 600 bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02ai_chain_abab.txt
