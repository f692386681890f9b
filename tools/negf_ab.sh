for cfg in "4 3" "2 3" "2 6" "1 6" "8 3" "4 2"; do set -- $cfg; echo "batch=$1 slots=$2"; SCLMD_BPT_BATCH=$1 SCLMD_BPT_SLOTS=$2 python tools/probe_bpt.py 12501 3 2>/dev/null | tail -2; done
