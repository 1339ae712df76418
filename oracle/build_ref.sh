#!/bin/bash
# oracle/build_ref.sh -- builds oracle/_ref/libsdf_ref.so from the reference's OWN hot-path text.
# The text (sdf_generation.hpp:19-271) is sed-extracted from /root/reference where it lies, written only into
# the git-ignored oracle/_ref/, and compiled against oracle/ref_shim.hpp.  No reference source is committed.
# On the GPU box /root/reference is absent: the prebuilt .so (shipped by gpurun) is used as is.
set -e
cd "$(dirname "$0")"
REF=${SDF_REFERENCE_ROOT:-/root/reference}
SRC="$REF/include/sdf_tools/sdf_generation.hpp"
mkdir -p _ref
if [ -f "$SRC" ]; then
    sed -n '19,271p' "$SRC" > _ref/ref_sdf_generation.inc
    g++ -O3 -std=c++17 -fPIC -shared -o _ref/libsdf_ref.so ref_driver.cpp
    echo "built oracle/_ref/libsdf_ref.so from $SRC"
elif [ -f _ref/libsdf_ref.so ]; then
    echo "reference tree absent; keeping prebuilt oracle/_ref/libsdf_ref.so"
else
    echo "reference tree absent and no prebuilt oracle/_ref/libsdf_ref.so" >&2
    exit 3
fi
