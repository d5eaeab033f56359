python tools/precision_plan.py --weights random 2>&1 | tee gpurun_out/r02k_precision_plan_random.txt | head -70
python tools/precision_plan.py --weights semantic 2>&1 | tee gpurun_out/r02k_precision_plan_semantic.txt | head -70
python tools/bench_mixed_geometry.py 2>&1 | tee gpurun_out/r02k_mixed_geometry.txt
