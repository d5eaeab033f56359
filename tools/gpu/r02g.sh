timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "pair or dispatcher" 2>&1 | tail -8
run() { echo "== $*"; env "$@" python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], {k: round(v['ms_per_page'],3) for k,v in d['roofline']['groups'].items()})"; }
for i in 1 2 3; do
  run SBB_PAIR=0
  run SBB_PAIR=1 SBB_PAIR_HEAD=0
  run SBB_PAIR=1 SBB_PAIR_HEAD=1
done 2>&1 | tee gpurun_out/r02g_pair_bench_abab.txt
