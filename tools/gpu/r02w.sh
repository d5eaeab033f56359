TAG=r02z
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_pair_kernel<.{0,8}0, .{0,8}128>' -s 87 -c 1 -f -o gpurun_out/${TAG}_dec2_pair python tools/prof_page.py --pages 2 > gpurun_out/${TAG}_ncu_dec2.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_dec2.log
timeout 400 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_cfg2.json; cut -c1-300 gpurun_out/${TAG}_bench_cfg2.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --config 3 --steps 24 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_cfg3.json; cut -c1-300 gpurun_out/${TAG}_bench_cfg3.json
tail -3 gpurun_out/${TAG}_bench.err
