#!/bin/bash
# usage: tools/build_base.sh [rev]  — builds the library of a git revision (default HEAD) into tools/exp/libti5_base.so,
# for same-box A/B runs against the working tree (TI5_LIB=$PWD/tools/exp/libti5_base.so python bench.py ...)
rev=${1:-HEAD}
rm -rf /tmp/ti5_base && mkdir -p /tmp/ti5_base tools/exp
git archive $rev ti5_isaacgym_b200/csrc include | tar -x -C /tmp/ti5_base
cd /tmp/ti5_base/ti5_isaacgym_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared -o $OLDPWD/tools/exp/libti5_base.so ti5_api.cu ti5_substep.cu ti5_post_physics.cu ti5_observe.cu ti5_heights.cu ti5_gae.cu ti5_rollout.cu && echo built base $rev
