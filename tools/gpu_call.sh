#!/bin/bash
# usage: tools/gpu_call.sh <script-file> <logfile> [timeout]  -- retries while the pod answers "busy" (rc 3 / transient)
script="$1"; log="$2"; to="${3:-1500}"
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$(cat "$script")" > "$log" 2>&1
  if grep -q "status=transient\|no box\|retry in a few minutes" "$log"; then sleep 90; continue; fi
  break
done
echo "gpu_call finished (attempt $attempt)" >> "$log"
