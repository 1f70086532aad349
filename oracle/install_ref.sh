#!/usr/bin/env bash
# Assembles the runtime of the reference build ($1 = build tree) in $2 (oracle/_ref), mirroring the
# layout of the build tree (which is what the reference's own CI runs from, ci.yml:15):
#     _ref/lib*.so            libmitsuba, libdrjit-core/-extra, libnanothread, image libraries
#     _ref/plugins/*.so       the plugins (resolved next to libmitsuba.so, MI/src/core/plugin.cpp:93)
#     _ref/python/{drjit,mitsuba}
# The build-tree rpaths are absolute (they name the build directory), so oracle/ref.py pre-loads
# the shared libraries by path before importing the modules.  Binaries are stripped; stubs dropped.
set -euo pipefail
BUILD="$1"; OUT="$2"
rm -rf "$OUT"
mkdir -p "$OUT/plugins" "$OUT/python"
cp "$BUILD"/*.so "$OUT/"
cp "$BUILD"/plugins/*.so "$OUT/plugins/"
cp -r "$BUILD/python/drjit" "$BUILD/python/mitsuba" "$OUT/python/"
find "$OUT/python" \( -name '*.pyi' -o -name 'py.typed' -o -name '__pycache__' \) -prune -exec rm -rf {} + 2>/dev/null || true
rm -rf "$OUT/python/mitsuba/mitsuba_stubs" "$OUT/python/mitsuba/python/test" 2>/dev/null || true
find "$OUT" -name '*.so' -type f -exec strip --strip-unneeded {} + 2>/dev/null || true
{ echo "built: $(date -u +%Y-%m-%dT%H:%M:%SZ)";
  echo "source: ${ERTB_REF_SRC:-/root/reference/ext/mitsuba}";
  grep "MI_DEFAULT_VARIANTS" "$BUILD/configure.log" | head -1;
  echo "ninja failed targets: $(grep -c FAILED "$BUILD/build.log" || true)";
  echo "size: $(du -sh "$OUT" | cut -f1)"; } > "$OUT/BUILD_INFO.txt"
cat "$OUT/BUILD_INFO.txt"
