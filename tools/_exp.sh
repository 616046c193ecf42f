for dbg in 0 1 2 3 4 5; do echo "SSPBO_DEBUG=$dbg"; SSPBO_DEBUG=$dbg python tools/kbench.py --ranks 10,60,127 --engines lr --reps 3 2>&1 | grep rank; done
