#!/bin/bash
# Build tuning variants of the library side by side (gpurun_out/lib_<name>.so is scratch, not shipped).
#   tools/variants.sh name "-DHPF_UNROLL_T=2 -DHPF_SWEEP_MINBLOCKS=3" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/../hgaprec_b200/csrc"
mkdir -p ../../build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( make -B OUT=../../build/variants/lib_$name.so EXTRA="$flags" LOG=/tmp/build_$name.log > /dev/null && \
    grep -A3 "12sweep_kernelILi8ELi4ELb0\|tile_sweep_kernelILi8ELi4ELb0" /tmp/build_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name" ) &
done
wait
