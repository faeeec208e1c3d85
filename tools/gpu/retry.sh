#!/bin/bash
# retry.sh <timeout> <script>: runs a GPU script through gpurun, retrying while the pod has no free slot
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $1 -- bash $2 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out" | tail -${3:-30}
  exit 0
done
echo "no GPU slot after 20 tries"
