#!/usr/bin/env bash
# link_shim_mains.sh -- TEST INFRASTRUCTURE.
# Links the reference's UNMODIFIED main.c / ekf.c objects (compiled by build_ref.sh from /root/reference
# into oracle/_ref/obj) against the product's per-shape shim: the reference's own driver program running
# on the B200 library (tests/test_shim_gpu.py).  Outputs stay in oracle/_ref/ (git-ignored, shipped to the
# GPU box).  Nothing of the product build depends on this script.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OBJ="$HERE/_ref/obj"
BUILD="$(cd "$HERE/.." && pwd)/nav-slam_b200/_build"
SHAPES="${SHAPES:-8x8 5x33 16x1800 64x2048}"
CC="${CC:-gcc}"
n=0
for shape in $SHAPES; do
    main_o="$OBJ/main_$shape.o"; nomain_o="$OBJ/main_nomain_$shape.o"; ekf_o="$OBJ/ekf_$shape.o"
    so="$BUILD/libnavslam_shim_$shape.so"
    if [ ! -f "$main_o" ] || [ ! -f "$so" ]; then continue; fi
    common="-L$BUILD -lnavslam_shim_$shape -lnavslam_b200 -Wl,-rpath,$BUILD -lm -l:libjansson.so.4"
    $CC "$main_o" "$ekf_o" $common -o "$HERE/_ref/navshim_main_$shape"
    $CC -O2 "$HERE/l9_main.c" "$nomain_o" "$ekf_o" $common -o "$HERE/_ref/navshim_l9_$shape"
    n=$((n + 2))
done
echo "link_shim_mains: $n programs"
