timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "pair" 2>&1 | tail -3
F='page 2800|res3b_branch2b|res4b_branch2b|res5b_branch2b|dec1|dec2|dec3|dec4|dec5|sum of' LIBS="libsbb_prev.so libsbb_textline.so libsbb_prev.so libsbb_textline.so" bash tools/exp_ab.sh 2>&1 | tee gpurun_out/r02p_arrive_scope_abab.txt
