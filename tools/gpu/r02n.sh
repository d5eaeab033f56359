SBB_SUBBATCH="3:4,4:2,5:2" timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "config2 or spanning" 2>&1 | tail -4
for v in "" "4:2" "4:2,5:2" "3:2,4:2" "3:4,4:2" "" "4:2" "4:3" "2:8,3:4,4:2"; do
  echo "== SBB_SUBBATCH=$v"; SBB_SUBBATCH="$v" python tools/gpu_diag.py --stage time --iters 8 2>&1 | grep -E "page 2800"
done 2>&1 | tee gpurun_out/r02n_subbatch.txt
