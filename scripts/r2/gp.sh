#!/bin/bash
# retry wrapper around gpurun: gp.sh <out-file> <timeout> [--gpus N] -- cmd   (retries while the pod answers busy / transient)
out=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $out 2>&1
  if grep -q "status=transient\|rc=3\|no box\|busy" $out && ! grep -q "status=ok" $out; then sleep 45; continue; fi
  break
done
tail -40 $out
