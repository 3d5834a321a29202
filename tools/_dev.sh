#!/bin/bash
# usage: tools/_dev.sh tag   — build quietly, run tools/_run.sh on the GPU box, print the log

cd /root/repo/deltarice_b200/csrc && make -j8 2>&1 | grep -E "error|spill|rror:" || true
cd /root/repo; gpurun --timeout 900 -- "bash tools/_run.sh > gpurun_out/$1.log 2>&1" | grep -E "status|left"
cat /root/repo/gpurun_out/$1.log
