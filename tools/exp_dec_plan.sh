# same-box A/B of the decoder plan knobs (SBB_DEC_RECT: tile shapes for the kept regions, SBB_DEC5_MERGED:
# dec5 as one merged-parity N = 128 GEMM): per-layer CUDA-event times of the 2800x2000 page
F='page 2800|conv1 |dec_v4|dec1|dec2|dec3|dec4|dec5|sum of|host-buffer'
for cfg in "SBB_DEC_RECT=0 SBB_DEC5_MERGED=0" "SBB_DEC_RECT=1 SBB_DEC5_MERGED=0" "SBB_DEC_RECT=0 SBB_DEC5_MERGED=1" "SBB_DEC_RECT=1 SBB_DEC5_MERGED=1"; do
  echo "== $cfg"; env $cfg python tools/gpu_diag.py --stage time --iters 5 2>&1 | grep -E "$F"
done
