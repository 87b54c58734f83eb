#!/bin/bash
# A/B builds of libf2d.so with different kernel switches (git-ignored, travel with gpurun): tools/build_variants.sh ps0 m0
set -e
cd "$(dirname "$0")/../fluid-2d_b200/csrc"
build() { name=$1; shift; make -j8 OUT="$PWD/../libf2d_$name.so" BUILD="$PWD/build_$name" VARIANT="$*" > /dev/null; echo "built libf2d_$name.so ($*)"; }
for v in "$@"; do
  case $v in
    ps0) build ps0 -DF2D_PRESSURE_SCALED=0 ;;   # pressure levels unscaled: 4 FADD + FMUL per cell-sweep
    m0)  build m0 -DF2D_RHS_MIRROR=0 ;;         # plain 16-slot rhs ring instead of the mirrored one
    ls0) build ls0 -DF2D_LEVEL_SPLIT=0 ;;       # all T levels of a row step as one dependent chain
    *) echo "unknown variant $v"; exit 1 ;;
  esac
done
