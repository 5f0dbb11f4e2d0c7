for w in 2 4 8 16; do echo "W=$w"; ZK_EXEC_WARPS=$w ZK_PROF_KIND=mix ZK_PROF_MB=1024 ZK_PROF_REPS=3 python tools/prof_decode.py 2>&1 | tail -1; done
