#!/bin/bash
# Builds alternative versions of the product library (experiments only, loaded through DLT_LIB_PATH) into build/variants/.
#   tools/build_variants.sh NAME "-DFLAG=1 ..." [NAME2 "..."]...
set -e
cd "$(dirname "$0")/.."
CSRC=dxt_lossless_transform_b200/csrc
mkdir -p build/variants
FLAGS=$(grep -v '^#' $CSRC/NVCC_FLAGS.txt | tr '\n' ' ')
SRCS=$(grep -v '^#' $CSRC/SOURCES.txt | sed "s|^|$CSRC/|" | tr '\n' ' ')
while [ $# -ge 2 ]; do
  echo "building build/variants/lib$1.so with: $2"
  nvcc $FLAGS $2 -shared -o build/variants/lib$1.so $SRCS &
  shift 2
done
wait
ls -la build/variants
