# same-box A/B: alternative builds of the library (SBB_LIB) vs the current build, per-layer CUDA-event times
F=${F:-'page 2800|conv1 |res2a_branch2a|res2a_branch2b|res2b_branch2a|res2b_branch2c|res3b_branch2b|res4b_branch2b|res4b_branch2c|res3a_branch2c|res5b_branch2b|dec_v4|dec1|dec2|dec3|dec4|dec5|sum of'}
for lib in ${LIBS:-libsbb_prev.so libsbb_textline.so}; do
  echo "== $lib"; SBB_LIB=$PWD/sbb_textline_detection_b200/$lib python tools/gpu_diag.py --stage time --iters 5 2>&1 | grep -E "$F"
done
