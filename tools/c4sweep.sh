for n in 21 42 63 64; do for w in 1 0; do python tools/config4_timing.py $n 401 $w 2>&1 | tail -1; done; done
