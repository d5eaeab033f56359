python tools/mint_gpu_semantic_labels.py 2>&1 | tail -4
timeout 600 python -m pytest tests/test_semantic_run.py tests/test_gpu_parity.py::test_three_model_pipeline_vs_oracle -x -q -m gpu 2>&1 | tail -12
for v in 0 1 0 1 0 1; do
  echo "== SBB_IMG_BOXES=$v"
  SBB_IMG_BOXES=$v python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"
done 2>&1 | tee gpurun_out/r02c_img_boxes_bench_ababab.txt
