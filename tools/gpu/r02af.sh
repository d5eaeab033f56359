# what the weight (B) rows cost the operand path, in SM CYCLES per item (clock-independent: the power cap moves with the data):
# role profile of the pair kernel with and without the B loads (experiment builds; the no-B results are wrong by construction)
for lib in libsbb_roles.so libsbb_roles_nob.so; do
  echo "== $lib"
  SBB_LIB=$PWD/sbb_textline_detection_b200/$lib SBB_DEBUG=16 python tools/gpu_diag.py --stage time --iters 1 2>&1 | grep -E "roles" | head -58
done | tee gpurun_out/r02af_no_b_rows_cycles.txt
