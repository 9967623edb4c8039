#!/bin/bash
# Builds the drop-in driver: the reference's own objects (compiled by oracle/build_ref.sh from the sources
# under $MPM_REFERENCE, minus its main.o) + GpuTasks.cpp (this repo) + libmpmgpu.so  ->  host/_build/NairnMPM_gpu
set -e
R=${MPM_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OBJ=$ROOT/oracle/_ref/obj
[ -d "$R" ] || { echo "build_host: no reference at $R (prebuilt host/_build is used on the GPU box)"; exit 0; }
[ -d "$OBJ" ] || bash "$ROOT/oracle/build_ref.sh"
mkdir -p "$HERE/_build"
g++ -c -O2 -fopenmp -std=c++11 -fPIC -w -I"$R/NairnMPM/src" -I"$R/Common/Headers" -I"$R/Common" -I"$ROOT/oracle/xerces_shim" \
    -include "$R/NairnMPM/src/System/MPMPrefix.hpp" "$HERE/GpuTasks.cpp" -o "$HERE/_build/GpuTasks.o"
g++ -fopenmp -o "$HERE/_build/NairnMPM_gpu" "$HERE/_build/GpuTasks.o" \
    $(ls "$OBJ"/*.o | grep -v -e '/main.o$' -e '/ref_harness.o$') \
    -L"$ROOT/nairn_mpm_fea_b200" -lmpmgpu -lexpat -Wl,-rpath,'$ORIGIN/../..'
echo "build_host: ok -> $HERE/_build/NairnMPM_gpu"
