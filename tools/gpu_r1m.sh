#!/bin/bash
for pr in fp16 split split_act; do
  echo "== $pr"; timeout 300 python tools/profile_step.py --arch resnet --blocks 10 --precision $pr --playouts 40 --games 4096 2>&1 | tr '\n' ' ' | sed 's/phase ms.*//'; echo
done
