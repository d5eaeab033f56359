# same-box A/B of an environment knob of the CURRENT library: per-layer CUDA-event times of one 2800x2000 page
#   VAR=SBB_IMG_BOXES VALS="0 1 0 1" bash tools/exp_env_ab.sh
F=${F:-'page 2800|conv1 |res2a_branch2b|res3a_branch2a|res3a_branch2c|res3b_branch2b|res4a_branch2c|res4b_branch2b|res5a_branch2c|res5b_branch2b|dec2|sum of'}
for v in ${VALS:-0 1 0 1}; do
  echo "== $VAR=$v"; env $VAR=$v python tools/gpu_diag.py --stage time --iters 5 2>&1 | grep -E "$F"
done
