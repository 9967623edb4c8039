#!/bin/bash
# TEST INFRASTRUCTURE -- builds the CHECKER, never the product.
#
# Compiles the reference NairnMPM sources, unmodified and where they lie under
# $MPM_REFERENCE (default /root/reference), into oracle/_ref/:
#   oracle/_ref/NairnMPM              the reference's own CLI (used as the "reference" CPU baseline)
#   oracle/_ref/libnairnmpm_ref.so    the same objects + oracle/ref_harness.cpp (step-by-step C access
#                                     to the reference's particle and node state, for parity tests)
# The reference's own build system is NOT run: the object list and name->path map are read from
# NairnMPM/build/makefile (objects = ... at :550-581, "name = $(src|com)/path" lines at :167-362) and
# each TU is compiled directly with the reference's flags (-O3 -fopenmp -std=c++11, makefile:129) plus
# -fPIC -w, force-including MPMPrefix.hpp (makefile:365).  Xerces-C (not installed) is replaced by the
# expat-backed stand-in in nairn_mpm_fea_b200/host/xerces_shim (shared with the drop-in driver's build).  Nothing is copied out of the reference tree.
set -e
R=${MPM_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
MK=$R/NairnMPM/build/makefile
[ -f "$MK" ] || { echo "build_ref: no reference at $R (fine on the GPU box: prebuilt oracle/_ref is used)"; exit 0; }
mkdir -p "$OUT/obj"
JOBS=${JOBS:-$(nproc)}
OPT=${REF_OPT:--O3}

OBJS=$(awk '/^objects/,/^$/' "$MK" | tr -d '\\' | sed 's/objects =//' | tr -s ' \t\n' ' ')

compile_one() {
    o=$1; b=${o%.o}
    p=$(grep -E "^$b ?= ?" "$MK" | head -1 | sed 's/.*= *//' | sed "s#\$(src)#$R/NairnMPM/src#; s#\$(com)#$R/Common#")
    [ -z "$p" ] && { echo "build_ref: no path for $b"; return 1; }
    [ "$OUT/obj/$o" -nt "$p.cpp" ] && return 0
    g++ -c $OPT -fopenmp -std=c++11 -fPIC -w \
        -I"$R/NairnMPM/src" -I"$R/Common/Headers" -I"$R/Common" -I"$HERE/../nairn_mpm_fea_b200/host/xerces_shim" \
        -include "$R/NairnMPM/src/System/MPMPrefix.hpp" "$p.cpp" -o "$OUT/obj/$o" \
        || { echo "build_ref: FAILED $b"; return 1; }
}
export -f compile_one; export R HERE OUT MK OPT
echo $OBJS | tr ' ' '\n' | grep -v '^$' | xargs -P "$JOBS" -I{} bash -c 'compile_one {}'

# the reference CLI
g++ -fopenmp -o "$OUT/NairnMPM" $(for o in $OBJS; do echo "$OUT/obj/$o"; done) -lexpat

# harness library = reference objects (minus main.o) + our C accessors
g++ -c $OPT -fopenmp -std=c++11 -fPIC -w \
    -I"$R/NairnMPM/src" -I"$R/Common/Headers" -I"$R/Common" -I"$HERE/../nairn_mpm_fea_b200/host/xerces_shim" \
    -include "$R/NairnMPM/src/System/MPMPrefix.hpp" "$HERE/ref_harness.cpp" -o "$OUT/obj/ref_harness.o"
g++ -shared -fopenmp -o "$OUT/libnairnmpm_ref.so" "$OUT/obj/ref_harness.o" \
    $(for o in $OBJS; do [ "$o" = main.o ] || echo "$OUT/obj/$o"; done) -lexpat
echo "build_ref: ok -> $OUT/NairnMPM, $OUT/libnairnmpm_ref.so"
