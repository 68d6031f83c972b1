#!/bin/bash
# A/B builds of the library with different -D switches for one source file: tools/_variants/<name>.so
# usage: tools/build_variants.sh k4_ppo_lag "name1:-DFOO=1" "name2:-DFOO=2 -DBAR" ...
set -e
cd "$(dirname "$0")/.."
SRC=$1; shift
mkdir -p tools/_variants
python -m icrl_b200.build > /dev/null
for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC $flags \
        -c icrl_b200/csrc/$SRC.cu -o tools/_variants/$SRC.$name.o
    objs=$(for f in icrl_b200/csrc/*.cu; do b=$(basename $f .cu); [ "$b" != "$SRC" ] && echo icrl_b200/build/$b.o; done)
    nvcc -shared -o tools/_variants/$name.so $objs tools/_variants/$SRC.$name.o -lcudart
    rm tools/_variants/$SRC.$name.o
    echo "built tools/_variants/$name.so ($flags)"
done
