for so in variants/enc_v11_w4.so variants/enc_v11_w2.so variants/enc_v11_w1.so variants/enc_v10_w4.so variants/enc_v10_w2.so; do echo -n "$so : "; LZ4B200_SO_OVERRIDE=$PWD/$so python bench.py --steps 3 --warmup 3 --quick 2>&1 | tail -1; done
LZ4B200_SO_OVERRIDE=$PWD/variants/enc_v11_w1.so python tests/gpu_quick.py 2>&1 | tail -1
