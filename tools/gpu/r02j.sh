SBB_DEBUG=16 python tools/gpu_diag.py --stage time --iters 1 2>&1 | grep -E "roles|pair" | head -70 | tee gpurun_out/r02j_role_wait_cycles.txt
