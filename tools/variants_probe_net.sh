# block-size / occupancy variants of the network step kernels (built with -DHWY_NET_BLOCK_THREADS / -DHWY_NET_MIN_BLOCKS)
export QB_CONFIGS=cfg3,cfg4
V=$PWD/highwayenv_b200/csrc/variants
python tools/quick_bench.py base_256x1
for v in n128_4 n512_1 n256_3 n192_2; do
  HWYB200_LIB=$V/libhwyb200_$v.so python tools/quick_bench.py $v
done
