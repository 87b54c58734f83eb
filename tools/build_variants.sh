#!/bin/bash
# A/B builds of libf2d.so with different kernel switches (git-ignored, travel with gpurun).
set -e
cd "$(dirname "$0")/../fluid-2d_b200/csrc"
build() { name=$1; shift; make -j8 OUT="$PWD/../libf2d_$name.so" BUILD="$PWD/build_$name" VARIANT="$*" > /dev/null; echo "built libf2d_$name.so ($*)"; }
for v in "$@"; do
  case $v in
    g0m1) build g0m1 -DF2D_RHS_GEN=0 -DF2D_RHS_MIRROR=1 ;;   # round-2 v3 kernel: one rhs LDS.128 per level and step
    g1m1) build g1m1 -DF2D_RHS_GEN=1 -DF2D_RHS_MIRROR=1 ;;
    g1m0) build g1m0 -DF2D_RHS_GEN=1 -DF2D_RHS_MIRROR=0 ;;
    ps0)  build ps0 -DF2D_PRESSURE_SCALED=0 ;;                # pressure levels unscaled: 4 FADD + FMUL per cell-sweep
    *) echo "unknown variant $v"; exit 1 ;;
  esac
done
