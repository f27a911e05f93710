#!/usr/bin/env bash
# TEST INFRASTRUCTURE - builds the reference's own C implementation of the BA path.
#
# Compiles Lourakis' sba 1.6 + libsbaprojs *from the sources where they lie* inside
# /root/reference/argus_gui/resources/sba_lib_sources.tar.gz (the only place the native hot path
# exists in the reference tree).  Nothing is copied into tracked files:
# the tarball is unpacked into oracle/_ref/src (git-ignored) and the outputs are
#   oracle/_ref/libsba.so, oracle/_ref/libsbaprojs.so, oracle/_ref/eucsbademo, oracle/_ref/demo/*.txt
# The reference's own Makefile/CMake is not run; this is the short recipe of SURVEY.md appendix C with
# the reference's flags (-O3 -funroll-loops, sba-1.6p/Makefile:10).
# LAPACK comes from the OpenBLAS bundled in the opencv_python_headless wheel (the image has no system
# LAPACK and no gfortran).
#
# oracle/_ref/ is git-ignored but NOT gpurun-ignored, so the built files travel to the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${ARGUS_REFERENCE:-/root/reference}"
TARBALL="$REF/argus_gui/resources/sba_lib_sources.tar.gz"
OUT="$HERE/_ref"
if [ ! -f "$TARBALL" ]; then
  if [ -f "$OUT/libsba.so" ] && [ -f "$OUT/libsbaprojs.so" ]; then
    echo "oracle/_ref: reference tree absent, using prebuilt files"; exit 0
  fi
  echo "oracle/_ref: reference tarball not found at $TARBALL and no prebuilt files" >&2; exit 1
fi
PY="${PYTHON:-python}"
BLASDIR="$($PY - <<'PYEOF'
import os, glob, importlib.util
spec = importlib.util.find_spec("cv2")
base = os.path.dirname(os.path.dirname(spec.origin))
for d in ("opencv_python_headless.libs", "opencv_python.libs", "opencv_contrib_python.libs"):
    p = os.path.join(base, d)
    if glob.glob(os.path.join(p, "libopenblas*.so*")):
        print(p); break
PYEOF
)"
[ -n "$BLASDIR" ] || { echo "no OpenBLAS found next to cv2" >&2; exit 1; }
BLAS="$(basename "$(ls "$BLASDIR"/libopenblas*.so* | head -1)")"
mkdir -p "$OUT/src" "$OUT/demo"
tar xzf "$TARBALL" -C "$OUT/src"
S="$OUT/src/sba-1.6p"
rm -f "$S"/*.gch "$S"/libsbaprojs/*.gch
CFLAGS="-O3 -funroll-loops -fPIC -w"
B="$OUT/build"; mkdir -p "$B"
for f in sba_levmar sba_levmar_wrap sba_lapack sba_crsm sba_chkjac; do
  gcc $CFLAGS -I"$S" -c "$S/$f.c" -o "$B/$f.o"
done
gcc -shared -o "$OUT/libsba.so" "$B"/sba_levmar.o "$B"/sba_levmar_wrap.o "$B"/sba_lapack.o "$B"/sba_crsm.o "$B"/sba_chkjac.o \
    -L"$BLASDIR" -l:"$BLAS" -Wl,-rpath,"$BLASDIR" -lm
gcc $CFLAGS -I"$S" -I"$S/libsbaprojs" -c "$S/libsbaprojs/sbaprojs.c" -o "$B/sbaprojs.o"
gcc $CFLAGS -I"$S" -I"$S/libsbaprojs" -c "$S/libsbaprojs/imgproj.c" -o "$B/imgproj.o"
gcc -shared -o "$OUT/libsbaprojs.so" "$B/sbaprojs.o" "$B/imgproj.o" -L"$OUT" -lsba -Wl,-rpath,'$ORIGIN' -lm
gcc -O3 -funroll-loops -w -I"$S" -I"$S/demo" -o "$OUT/eucsbademo" "$S/demo/eucsbademo.c" "$S/demo/imgproj.c" "$S/demo/readparams.c" \
    -L"$OUT" -lsba -Wl,-rpath,'$ORIGIN' -L"$BLASDIR" -l:"$BLAS" -Wl,-rpath,"$BLASDIR" -lm
# the reference's own demo data sets + README known answers (golden inputs)
cp "$S"/demo/*.txt "$OUT/demo/"
echo "$BLASDIR/$BLAS" > "$OUT/blas_path.txt"
rm -rf "$B"
echo "oracle/_ref built: $(ls "$OUT")"
