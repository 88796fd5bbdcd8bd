# usage: ab.sh name1 name2 ... — quick_bench of variant libraries (multiviewstitch_b200/variants/libmvs_<name>.so); "base" = the built library
for v in "$@"; do
  if [ "$v" == "base" ]; then bash scripts/quick_bench.sh $v; else MVS_B200_LIB=$PWD/multiviewstitch_b200/variants/libmvs_$v.so bash scripts/quick_bench.sh $v; fi
done
