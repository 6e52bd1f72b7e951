#!/bin/bash
# usage: tools/sass_loop_stats.sh <mangled kernel name>  -- opcode histogram of the channel loop (first LDS.128 .. loop end)
f=${1:-_ZN4bflk15das_tile_kernelILi6ELi16ELb1ELb1EEEvNS_10KernelArgsE}
cuobjdump -sass -fun "$f" beamforming-lk_b200/build/das_tile.o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+//; s/\s*\/\*.*$//' > /tmp/k.txt
# the first copy of the loop: from the 5th LDS.128 (after the preamble's entry loads) to the first BSYNC.RECONVERGENT B1
awk '/LDS.128/{c++} c>=5' /tmp/k.txt | awk '/BSYNC.RECONVERGENT B1/{exit} {print}' > /tmp/loop.txt
wc -l < /tmp/loop.txt
sed -E 's/^@!?U?P[0-9T]+ +//' /tmp/loop.txt | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -14
