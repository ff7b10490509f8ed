#!/bin/bash
# compute-sanitizer directly on the unmodified host + shim (a child started from python is not instrumented):
#   tools/probes/sanitize_host.sh <tool> <WxH> <frames> [extra sanitizer args]
set -e
cd "$(dirname "$0")/../.."
ROOT=$PWD
tool=$1; size=$2; frames=$3; shift 3
w=${size%x*}; h=${size#*x}
tmp=$(mktemp -d /dev/shm/sanit_XXXX)
python - <<PY
import sys
sys.path.insert(0, "$ROOT/tools"); sys.path.insert(0, "$ROOT/tests")
import gen_y4m, _trace
gen_y4m.write_y4m("$tmp/clip.y4m", $w, $h, $frames)
open("$tmp/GPU_kernels.cl", "w").write(_trace.GPU_STUB)
open("$tmp/CPU_kernels.cl", "w").write(_trace.CPU_STUB)
PY
cd $tmp
export LD_LIBRARY_PATH=$ROOT/vp8oclenc_b200/lib:$LD_LIBRARY_PATH
export VP8B200_HOST_PROFILE=reference
compute-sanitizer --tool $tool "$@" $ROOT/vp8oclenc_b200/bin/vp8enc -i clip.y4m -o out.ivf -qmin 24 -qmax 24 -g 150 -altref-range 5 -partitions 8 -threads 12 2>&1 | tail -40
md5sum out.ivf
rm -rf $tmp
