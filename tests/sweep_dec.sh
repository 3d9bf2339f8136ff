for so in variants/dec_defer0.so variants/dec_defer1.so; do for G in 16 32 8; do echo -n "$so G=$G : "; LZ4B200_SO_OVERRIDE=$PWD/$so LZ4B200_DEC_GROUP=$G python bench.py --steps 3 --warmup 3 --quick 2>&1 | tail -1; done; done
LZ4B200_SO_OVERRIDE=$PWD/variants/dec_defer1.so python tests/gpu_quick.py 2>&1 | tail -1
LZ4B200_SO_OVERRIDE=$PWD/variants/dec_defer1.so LZ4B200_DEC_GROUP=32 python tests/gpu_quick.py 2>&1 | tail -1
