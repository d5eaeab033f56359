export SBB_PAIR=1
for v in 8 4 2 16 8 4 2 1; do
  echo "== SBB_PAIR_MIN_CHUNKS=$v"; SBB_PAIR_MIN_CHUNKS=$v python tools/gpu_diag.py --stage time --iters 8 2>&1 | grep -E "page 2800|sum of"
done 2>&1 | tee gpurun_out/r02e_pair_min_chunks.txt
SBB_PAIR_MIN_CHUNKS=2 python tools/gpu_diag.py --stage time --iters 5 2>&1 | grep -E "branch2c|branch2a|dec_v" | head -40 | tee -a gpurun_out/r02e_pair_min_chunks.txt
