#!/bin/bash
# Builds the drop-in driver host/_build/NairnMPM_gpu = the reference driver with its step tasks bound to libmpmgpu:
#   the reference's own translation units, compiled unmodified where they lie under $MPM_REFERENCE (default
#   /root/reference) into host/_build/obj (object list and name->path map read from NairnMPM/build/makefile, the
#   reference's flags -O3 -fopenmp -std=c++11; its main.o is left out),
#   + GpuTasks.cpp (this repo: the MPMTask subclasses and the install hook)
#   + libmpmgpu.so
# Xerces-C (not installed here) is replaced by the expat-backed SAX2 stand-in in host/xerces_shim.
# Nothing under oracle/ is used: this is the product's reference-side binding, not the checker.
set -e
R=${MPM_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
MK=$R/NairnMPM/build/makefile
[ -f "$MK" ] || { echo "build_host: no reference at $R (prebuilt host/_build is used on the GPU box)"; exit 0; }
OBJ=$HERE/_build/obj
mkdir -p "$OBJ"
JOBS=${JOBS:-$(nproc)}
OBJS=$(awk '/^objects/,/^$/' "$MK" | tr -d '\\' | sed 's/objects =//' | tr -s ' \t\n' ' ')
INC="-I$R/NairnMPM/src -I$R/Common/Headers -I$R/Common -I$HERE/xerces_shim -include $R/NairnMPM/src/System/MPMPrefix.hpp"
compile_one() {
    o=$1; b=${o%.o}
    p=$(grep -E "^$b ?= ?" "$MK" | head -1 | sed 's/.*= *//' | sed "s#\$(src)#$R/NairnMPM/src#; s#\$(com)#$R/Common#")
    [ -z "$p" ] && { echo "build_host: no path for $b"; return 1; }
    [ "$OBJ/$o" -nt "$p.cpp" ] && return 0
    g++ -c -O3 -fopenmp -std=c++11 -fPIC -w $INC "$p.cpp" -o "$OBJ/$o" || { echo "build_host: FAILED $b"; return 1; }
}
export -f compile_one; export R HERE OBJ MK INC
echo $OBJS | tr ' ' '\n' | grep -v -e '^$' -e '^main.o$' | xargs -P "$JOBS" -I{} bash -c 'compile_one {}'
g++ -c -O2 -fopenmp -std=c++11 -fPIC -w $INC "$HERE/GpuTasks.cpp" -o "$HERE/_build/GpuTasks.o"
g++ -fopenmp -o "$HERE/_build/NairnMPM_gpu" "$HERE/_build/GpuTasks.o" \
    $(for o in $OBJS; do [ "$o" = main.o ] || echo "$OBJ/$o"; done) \
    -L"$ROOT/nairn_mpm_fea_b200" -lmpmgpu -lexpat -Wl,-rpath,'$ORIGIN/../..'
echo "build_host: ok -> $HERE/_build/NairnMPM_gpu"
