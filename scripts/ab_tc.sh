#!/bin/bash
# A/B of two builds of the library on ONE box (box-to-box spread is several per cent): bash scripts/ab_tc.sh <other .so> [ranks]
OTHER=$1
RANKS=${2:-"8,32,64,128,256"}
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
for rep in 1 2; do
  echo "--- this tree (rep $rep)"; python scripts/tc_time.py $RANKS 4096 1024 2>&1 | grep rank
  echo "--- $OTHER (rep $rep)"; SVDLSTM_LIB=$OTHER python scripts/tc_time.py $RANKS 4096 1024 2>&1 | grep rank
done
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
