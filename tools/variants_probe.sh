export QB_CONFIGS=cfg2,cfg1,cfg5
V=$PWD/highwayenv_b200/csrc/variants
python tools/quick_bench.py base_128regs
HWYB200_LIB=$V/libhwyb200_256_3.so python tools/quick_bench.py r80_256x3
HWYB200_LIB=$V/libhwyb200_256_4.so python tools/quick_bench.py r64_256x4
HWYB200_LIB=$V/libhwyb200_384_2.so python tools/quick_bench.py r80_192x4
HWYB200_LIB=$V/libhwyb200_384_2.so HWYB200_EPB=6 python tools/quick_bench.py r80_384x2
HWYB200_LIB=$V/libhwyb200_320_2.so HWYB200_EPB=5 python tools/quick_bench.py r96_320x2
HWYB200_LIB=$V/libhwyb200_320_2.so python tools/quick_bench.py r96_128x5
HWYB200_EPB=2 python tools/quick_bench.py base_epb2
HWYB200_EPB=8 python tools/quick_bench.py base_epb8
