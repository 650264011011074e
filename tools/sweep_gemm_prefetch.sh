for P in 0 4 8 16; do echo "== prefetch $P"; MOLCLR_GEMM_PREFETCH=$P CASE=fwd timeout 120 python tools/bench_gemm.py 2>&1 | grep -E "fwd1 x1   |fwd1 x3 derive|mixed|fwd2 x3 derive|fwd2 x1   "; done
