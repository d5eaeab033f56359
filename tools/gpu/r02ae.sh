# role wait-cycle profile of the CTA-pair kernel (experiment build -DSBB_X_ROLES): one page, every launch
SBB_LIB=$PWD/sbb_textline_detection_b200/libsbb_roles.so SBB_DEBUG=16 python tools/gpu_diag.py --stage time --iters 1 2>&1 | grep -E "roles|page 2800" | head -70 | tee gpurun_out/r02ae_role_wait_cycles_pair.txt
