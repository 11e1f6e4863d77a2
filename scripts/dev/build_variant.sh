#!/bin/bash
# Build an experiment variant of the product library next to the default one:
#   scripts/dev/build_variant.sh NAME -DMBAVO_PIPELINE=1 ...   ->  mba-vo_b200/lib/libmbavo_NAME.so  (select it with MBAVO_LIBRARY=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../../mba-vo_b200"
make -j8 BUILD=build_$name LIB=lib/libmbavo_$name.so VARIANT="$*" > /tmp/build_$name.log 2>&1 || { tail -20 /tmp/build_$name.log; exit 1; }
grep -c "spill" /tmp/build_$name.log || true
ls -la lib/libmbavo_$name.so
