timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "pair" 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
VAR=SBB_DEC4_MERGED VALS="0 1 0 1" F='page 2800|dec3|dec4|dec5|sum of' timeout 400 bash tools/exp_env_ab.sh 2>&1 | tee gpurun_out/r02h_dec4_merged_abab.txt
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], {k: round(v['ms_per_page'],3) for k,v in d['roofline']['groups'].items()})"; }
for i in 1 2; do
  run SBB_DEC4_MERGED=0
  run SBB_DEC4_MERGED=1
done 2>&1 | tee -a gpurun_out/r02h_dec4_merged_abab.txt
