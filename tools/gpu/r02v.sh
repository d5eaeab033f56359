timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -k "pair or every_layer or tile_logits or golden_tile" 2>&1 | tail -4
VAR=SBB_PAIR64 VALS="1 2 1 2" F='page 2800|conv1 |sum of' timeout 400 bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02v_stem_wide_abab.txt
