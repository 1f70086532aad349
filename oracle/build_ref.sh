#!/usr/bin/env bash
# Builds the REFERENCE kernel itself (Eradiate's vendored Mitsuba 3 fork + Dr.Jit + the Eradiate
# plugins, /root/reference/ext/mitsuba) out of tree and installs the runtime into oracle/_ref/.
#
# Test infrastructure only: oracle/_ref is the checker (reference renders that pin the C oracle
# port, the traverse key set, and the CPU arm of bench.py).  Nothing under eradiate_b200/ loads it.
# No reference SOURCE is copied into the repo: sources are compiled where they lie, the build tree
# lives under $ERTB_REF_BUILD (default /tmp/mi_build), only binaries + the generated python package
# land in oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it travels to the GPU box).
#
# Variants: scalar_mono (CPU arm, what Eradiate's "mono" mode runs, _mode.py:56-123),
# scalar_mono_double (parity reference), scalar_mono_polarized_double (Stokes parity reference),
# llvm_mono (north-star CPU arm; compiles without LLVM, needs a libLLVM at run time).
# A scalar-only variant list fails at configure (duplicate nanothread subdir), SURVEY 7-0.
# scalar_rgb must be in the list (libmitsuba's Resampler references its ReconstructionFilter), and a
# JIT variant without an `ad_` one leaves libmitsuba with undefined ad_* symbols unless drjit-extra
# is linked: -DMI_ENABLE_AUTODIFF=ON does just that (the macro is not used by any source file).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${ERTB_REF_SRC:-/root/reference/ext/mitsuba}"
BUILD="${ERTB_REF_BUILD:-/tmp/mi_build}"
OUT="$HERE/_ref"
JOBS="${ERTB_REF_JOBS:-$(nproc)}"
VARIANTS="scalar_rgb,scalar_mono,scalar_mono_double,scalar_mono_polarized_double,llvm_mono"

if [ ! -d "$SRC" ]; then
    echo "build_ref: $SRC not present (GPU box?) - using the prebuilt oracle/_ref as is" >&2
    exit 0
fi
mkdir -p "$BUILD" "$OUT"
# /opt/gcc/bin/g++ (first on PATH in this image) is a wrapper without lto-wrapper: the reference's
# CMake files switch LTO on for nanothread / drjit-core / nanobind, so name the real compiler.
cmake -S "$SRC" -B "$BUILD" -GNinja -DCMAKE_BUILD_TYPE=Release \
      -DCMAKE_C_COMPILER=/usr/bin/gcc -DCMAKE_CXX_COMPILER=/usr/bin/g++ \
      -DMI_DEFAULT_VARIANTS="$VARIANTS" -DMI_ENABLE_EMBREE=OFF -DMI_ENABLE_AUTODIFF=ON \
      -DCMAKE_POLICY_VERSION_MINIMUM=3.5 > "$BUILD/configure.log" 2>&1
# -k 0: the stub generators (*.pyi) import the freshly built module and may fail without harm
nice -n 10 ninja -C "$BUILD" -j"$JOBS" -k 0 > "$BUILD/build.log" 2>&1 || true
"$HERE/install_ref.sh" "$BUILD" "$OUT"
