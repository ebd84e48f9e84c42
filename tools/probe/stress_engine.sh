#!/bin/bash
# Repeats the engine parity tests until one fails (intermittent races); prints the assertion.
n=${1:-10}
sel=${2:-"uniform or keyframe_chain or edge"}
for i in $(seq 1 $n); do
  timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k "$sel" > /tmp/stress_out.txt 2>&1
  rc=$?
  if [ $rc -ne 0 ]; then
    echo "--- failed at iteration $i (rc $rc)"
    grep -E "^E  |Error|^tests/|^FAILED" /tmp/stress_out.txt | cut -c1-600 | head -30
    exit 1
  fi
done
echo "all $n iterations passed"
