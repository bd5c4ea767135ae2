#!/bin/bash
# build A/B variants of libssb200.so: tools/build_variants.sh name "flags" [name "flags" ...]
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  SSB_NVCC_FLAGS="$flags" python -c "from streamsculptor_b200 import _lib; _lib.build(out='build/variants/$name.so')" &
done
wait
ls -la build/variants
