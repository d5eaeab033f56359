# store-path experiment: what the TMA stores of the pair kernel cost per layer (x1: no stores, x2: no lo-plane store; WRONG results)
export F='page 2800|conv1 |res2a_branch2b|res2a_branch2c|res2b_branch2a|res2b_branch2c|res3b_branch2a|res3b_branch2b|res3b_branch2c|res4b_branch2a|res4b_branch2b|res4b_branch2c|res5b_branch2c|dec1|dec2|dec3|dec4|sum of'
LIBS="libsbb_textline.so libsbb_x1.so libsbb_x2.so libsbb_textline.so libsbb_x1.so libsbb_x2.so" bash tools/exp_ab.sh 2>&1 | tee gpurun_out/r02ab_store_path.txt
