#!/usr/bin/env bash
# build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's own C sources, where they lie under $REF
# (/root/reference, read-only), into oracle/_ref/ (git-ignored, shipped to the
# GPU box by gpurun).  No reference source is copied into the repository: the
# per-shape header / slam.c variants are produced with sed into a mktemp
# directory that is deleted when the script exits.
#
#   libnavref_<RxC>.so : slam.c + kdtree.c + pointcloud.c + ekf.c + main.c (main renamed) + oracle/ref_driver.c
#                        8x8 is built from the untouched sources.  Other shapes
#                        rewrite ONLY utils/pointcloud.h:9-10 (MAX_ROWS/MAX_COLS,
#                        unconditional #defines) and the two fixed
#                        [100] correspondence buffers at src/slam.c:214,301
#                         -- no arithmetic changes.
#   navref_main_8x8    : the unmodified program (main.c L5 handler, config 1)
#   navref_l9_<RxC>    : main.c with -Dmain=ref_main + oracle/l9_main.c (config 2)
#   *.o for the shim-linked mains are left in oracle/_ref/obj/ (main.c, ekf.c)
# Flags follow CMakeLists.txt:5,9 (C11 with GNU extensions, -O2).
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
SHAPES="${SHAPES:-8x8 5x33 16x1800 64x2048}"
CC="${CC:-gcc}"
CFLAGS="-std=gnu11 -O2 -fPIC -ffp-contract=off -w"

if [ ! -d "$REF/src" ]; then
    echo "build_ref: $REF not present; keeping prebuilt oracle/_ref as is" >&2
    exit 0
fi
mkdir -p "$OUT/obj"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

for shape in $SHAPES; do
    R="${shape%x*}"; C="${shape#*x}"
    st="$TMP/$shape"; mkdir -p "$st"
    inc="-I$REF/headers -I$REF/utils"
    slam="$REF/src/slam.c"
    pre=""
    if [ "$shape" != "8x8" ]; then
        sed -E "s/^#define MAX_ROWS .*/#define MAX_ROWS $R/; s/^#define MAX_COLS .*/#define MAX_COLS $C/" \
            "$REF/utils/pointcloud.h" > "$st/pointcloud.h"
        grep -q "#define MAX_ROWS $R\$" "$st/pointcloud.h"
        grep -q "#define MAX_COLS $C\$" "$st/pointcloud.h"
        sed -E 's/NeighborResult result\[100\];/static NeighborResult result[MAX_ROWS*MAX_COLS];/; s/double ErrDistance\[100\];/static double ErrDistance[MAX_ROWS*MAX_COLS];/' \
            "$REF/src/slam.c" > "$st/slam.c"
        grep -q 'static NeighborResult result\[MAX_ROWS\*MAX_COLS\];' "$st/slam.c"
        grep -q 'static double ErrDistance\[MAX_ROWS\*MAX_COLS\];' "$st/slam.c"
        slam="$st/slam.c"
        # the include guard POINTCLOUD_H makes every later #include "pointcloud.h" a no-op
        pre="-include $st/pointcloud.h"
    fi
    # objects of the caller side (main.c, ekf.c) for linking against the B200 shim
    $CC $CFLAGS $pre $inc -I"$HERE/jansson_compat" -c "$REF/src/main.c" -o "$OUT/obj/main_$shape.o"
    $CC $CFLAGS $pre $inc -Dmain=ref_main -I"$HERE/jansson_compat" -c "$REF/src/main.c" -o "$OUT/obj/main_nomain_$shape.o"
    $CC $CFLAGS $pre $inc -c "$REF/src/ekf.c" -o "$OUT/obj/ekf_$shape.o"
    # the library also carries main.c (as ref_main) so that its CSV reader L9_LidarProcessData is callable
    $CC $CFLAGS -shared $pre $inc "$slam" "$REF/utils/kdtree.c" "$REF/utils/pointcloud.c" \
        "$REF/src/ekf.c" "$OUT/obj/main_nomain_$shape.o" "$HERE/ref_driver.c" -lm -l:libjansson.so.4 \
        -o "$OUT/libnavref_$shape.so"
    # whole-program reference binaries
    $CC $CFLAGS $pre $inc "$OUT/obj/main_$shape.o" "$OUT/obj/ekf_$shape.o" "$slam" \
        "$REF/utils/kdtree.c" "$REF/utils/pointcloud.c" -lm -l:libjansson.so.4 -o "$OUT/navref_main_$shape"
    $CC $CFLAGS $pre $inc "$HERE/l9_main.c" "$OUT/obj/main_nomain_$shape.o" "$OUT/obj/ekf_$shape.o" "$slam" \
        "$REF/utils/kdtree.c" "$REF/utils/pointcloud.c" -lm -l:libjansson.so.4 -o "$OUT/navref_l9_$shape"
    echo "build_ref: $shape ok"
done
