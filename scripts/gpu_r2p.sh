#!/bin/bash
for v in 0 50 200 500; do echo "poll $v: $(LDP_EPI_POLL=$v timeout 300 python scripts/ops_profile.py poll$v 2>&1 | head -1)"; done
