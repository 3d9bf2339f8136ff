for cfg in "16 0 8" "16 0 10" "16 0 12" "16 0 16" "8 0 8" "8 0 10" "8 0 16" "32 0 12" "32 0 16"; do set -- $cfg; echo -n "G=$1 batched=$2 ctas=$3 : "; LZ4B200_DEC_GROUP=$1 LZ4B200_DEC_BATCHED=$2 LZ4B200_DEC_CTAS=$3 python bench.py --steps 3 --warmup 3 --quick 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print(round(d['decompress_ms'],3))
"; done
