#!/bin/bash
# quick validation: selected GPU tests, then tools/r2_env_ab.sh with the given env variants
#   TESTS="tests/a.py tests/b.py" bash tools/r2_quick.sh <tag> <variants...>
TAG=$1; shift
O=gpurun_out; mkdir -p $O
if [ -n "$TESTS" ]; then
  timeout -k 10 600 python -m pytest $TESTS -q -m gpu -x -p no:cacheprovider > $O/pytest_${TAG}.log 2>&1
  echo "tests rc=$?"; tail -12 $O/pytest_${TAG}.log | cut -c1-300
fi
bash tools/r2_env_ab.sh $TAG "$@"
