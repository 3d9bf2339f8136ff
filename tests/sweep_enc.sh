for G in 32 16 8; do echo "== ENC G=$G"; LZ4B200_ENC_GROUP=$G python tests/gpu_quick.py 2>&1 | tail -2; LZ4B200_ENC_GROUP=$G python bench.py --steps 3 --warmup 3 --quick 2>&1 | tail -1; done
