# One gpurun call: GPU tests, the ncu launch list of one 2800x2000 page (-> DRAM traffic per kernel group, which the
# bench line then carries), the bench line, and optionally full captures of two kernels.
#   bash tools/gpu_round_profile.sh r01l [full]
TAG=${1:-r01x}
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_launches_page2800x2000.csv python tools/prof_page.py --pages 2 > gpurun_out/${TAG}_prof_page.log 2>&1
python tools/ncu_launches.py gpurun_out/${TAG}_launches_page2800x2000.csv gpurun_out/${TAG}_prof_page.log > gpurun_out/${TAG}_launches_page2800x2000.txt; tail -8 gpurun_out/${TAG}_launches_page2800x2000.txt
python tools/ncu_traffic.py gpurun_out/${TAG}_launches_page2800x2000.csv gpurun_out/${TAG}_prof_page.log profiles/${TAG}_traffic.json > /dev/null && cp profiles/${TAG}_traffic.json gpurun_out/
timeout 300 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_1gpu.json; cat gpurun_out/${TAG}_bench_1gpu.json
if [ "$2" = full ]; then
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_tc_kernel<.{0,6}64, .{0,6}1, .{0,6}0>' -s 8 -c 1 -f -o gpurun_out/${TAG}_conv1 python tools/prof_page.py --pages 2 > gpurun_out/${TAG}_ncu_conv1.log 2>&1; tail -3 gpurun_out/${TAG}_ncu_conv1.log
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_gemm_tc_kernel<.{0,6}128, .{0,6}1, .{0,6}1>' -s 1 -c 1 -f -o gpurun_out/${TAG}_dec5 python tools/prof_page.py --pages 2 > gpurun_out/${TAG}_ncu_dec5.log 2>&1; tail -3 gpurun_out/${TAG}_ncu_dec5.log
fi
