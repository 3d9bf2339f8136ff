for KB in 100 132 164 196 228; do echo -n "smem_kb=$KB : "; LZ4B200_ENC_SMEM_KB=$KB python bench.py --steps 3 --warmup 3 --quick 2>&1 | tail -1; done
