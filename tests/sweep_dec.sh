for so in variants/var_a0_u0.so variants/var_a0_u1.so variants/var_a1_u0.so variants/var_a1_u1.so; do
for cfg in "16 16" "8 16" "32 16"; do set -- $cfg; echo -n "$so G=$1 ctas=$2 : "; LZ4B200_SO_OVERRIDE=$PWD/$so LZ4B200_DEC_GROUP=$1 LZ4B200_DEC_CTAS=$2 python bench.py --steps 3 --warmup 3 --quick 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print(round(d['decompress_ms'],3))
"; done; done
