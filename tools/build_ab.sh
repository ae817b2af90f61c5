#!/bin/bash
# build a variant of the library for A/B timing: tools/build_ab.sh <tag> [-DMACRO=... ...]  -> codeformer_b200/ab/lib_<tag>.so
tag=$1; shift
mkdir -p codeformer_b200/ab
CFB_BUILD_OUT=$PWD/codeformer_b200/ab/lib_$tag.so CFB_NVCC_EXTRA="$*" python -m codeformer_b200.build --force > /tmp/build_$tag.log 2>&1 && echo "built $tag" || { echo "FAILED $tag"; tail -20 /tmp/build_$tag.log; }
