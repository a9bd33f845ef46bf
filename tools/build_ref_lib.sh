#!/bin/sh
# Build libzvx.so of a given commit into tools/_ab/libzvx_<name>.so (A/B timing with tools/ab_interleaved.py).
# usage: tools/build_ref_lib.sh <commit> <name>
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$1" zerovox.cpp_b200/csrc include | tar -x -C "$TMP"
make -C "$TMP/zerovox.cpp_b200/csrc" -j8 >/dev/null 2>&1
mkdir -p "$ROOT/tools/_ab"
cp "$TMP/zerovox.cpp_b200/libzvx.so" "$ROOT/tools/_ab/libzvx_$2.so"
rm -rf "$TMP"
echo "tools/_ab/libzvx_$2.so"
