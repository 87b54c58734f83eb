#!/bin/bash
# A/B builds of libf2d.so with different kernel switches (git-ignored, travel with gpurun).
set -e
cd "$(dirname "$0")/../fluid-2d_b200/csrc"
build() { name=$1; shift; make -j8 OUT="$PWD/../libf2d_$name.so" BUILD="$PWD/build_$name" VARIANT="$*" > /dev/null; echo "built libf2d_$name.so ($*)"; }
build v1 -DF2D_RING_OR=1 -DF2D_SHFL_AHEAD=1
build v2 -DF2D_RING_OR=0 -DF2D_SHFL_AHEAD=1
build v3 -DF2D_RING_OR=0 -DF2D_SHFL_AHEAD=0
