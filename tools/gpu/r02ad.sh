# final-tree confirmation: GPU suite, smoke, default bench
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02ad_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/r02ad_tests.txt
timeout 400 python bench.py 2>gpurun_out/r02ad_bench.err | tail -1 > gpurun_out/r02ad_bench_cfg2.json; cut -c1-600 gpurun_out/r02ad_bench_cfg2.json; tail -3 gpurun_out/r02ad_bench.err
