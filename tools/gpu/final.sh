# One gpurun call for the round's final records:  bash tools/gpu/final.sh r02z
TAG=${1:-r02z}
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_tests.txt
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_launches_page2800x2000.csv python tools/prof_page.py --pages 2 > gpurun_out/${TAG}_prof_page.log 2>&1
python tools/ncu_launches.py gpurun_out/${TAG}_launches_page2800x2000.csv gpurun_out/${TAG}_prof_page.log > gpurun_out/${TAG}_launches_page2800x2000.txt; tail -8 gpurun_out/${TAG}_launches_page2800x2000.txt
python tools/ncu_traffic.py gpurun_out/${TAG}_launches_page2800x2000.csv gpurun_out/${TAG}_prof_page.log profiles/${TAG}_traffic.json > /dev/null && cp profiles/${TAG}_traffic.json gpurun_out/
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_pair_kernel<.{0,8}0, .{0,8}128, .{0,8}0>' -s 87 -c 1 -f -o gpurun_out/${TAG}_dec2_pair python tools/prof_page.py --pages 2 > gpurun_out/${TAG}_ncu_dec2.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_dec2.log
timeout 400 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_cfg2.json; cut -c1-400 gpurun_out/${TAG}_bench_cfg2.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_reference_arm.json; cut -c1-200 gpurun_out/${TAG}_bench_reference_arm.json
timeout 400 python bench.py --config 3 --steps 24 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_cfg3.json; cut -c1-300 gpurun_out/${TAG}_bench_cfg3.json
timeout 400 python bench.py --config 5 --steps 12 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_cfg5.json; cut -c1-300 gpurun_out/${TAG}_bench_cfg5.json
timeout 400 python bench.py --config 5o --steps 8 --no-cpu-baseline 2>>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_cfg5o.json; cut -c1-300 gpurun_out/${TAG}_bench_cfg5o.json
tail -3 gpurun_out/${TAG}_bench.err
