for v in "SBB_PAIR=1" "SBB_PAIR=2 SBB_PAIR_MIN_CHUNKS=8" "SBB_PAIR=2 SBB_PAIR_MIN_CHUNKS=4" "SBB_PAIR=2 SBB_PAIR_MIN_CHUNKS=2" "SBB_PAIR=1" "SBB_PAIR=2 SBB_PAIR_MIN_CHUNKS=8" "SBB_PAIR=2 SBB_PAIR_MIN_CHUNKS=4" "SBB_PAIR=2 SBB_PAIR_MIN_CHUNKS=2"; do
  echo "== $v"; env $v python tools/gpu_diag.py --stage time --iters 8 2>&1 | grep -E "page 2800|sum of"
done 2>&1 | tee gpurun_out/r02s_pair_scope_sweep.txt
env SBB_PAIR=2 SBB_PAIR_MIN_CHUNKS=2 python tools/gpu_diag.py --stage time --iters 5 2>&1 | grep -E "branch2c|branch2a|dec_v" | tee -a gpurun_out/r02s_pair_scope_sweep.txt
env SBB_PAIR=1 python tools/gpu_diag.py --stage time --iters 5 2>&1 | grep -E "branch2c|branch2a|dec_v" | tee -a gpurun_out/r02s_pair_scope_sweep.txt
