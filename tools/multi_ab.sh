# usage: bash tools/multi_ab.sh "libA.so libB.so ..." "n rounds world" ...   (same-box comparison of several builds)
libs=$1; shift
for w in "$@"; do
  for L in $libs; do
    timeout 150 python tools/ab_compare.py $L $L $w 2>&1 | head -1 | cut -c1-150
  done
done
