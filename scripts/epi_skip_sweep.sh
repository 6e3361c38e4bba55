for m in 0 1 2 4 8 16 31; do LDP_EPI_SKIP=$m python scripts/ops_profile.py r1g_skip$m > gpurun_out/ops_r1g_skip$m.log 2>&1; done; head -1 gpurun_out/ops_r1g_skip*.log
