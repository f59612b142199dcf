#!/bin/bash
for k in 4 3; do for t in 0 1; do echo "kernel $k tally $t"; python tools/ncu_one.py --kernel $k --tally $t --n-inner 1 --reps 4 | tail -2; done; done
