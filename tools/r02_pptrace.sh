GDRN_DW_TRACE=1 python tools/dw_bench.py 2>&1 | grep "pp trace" | sort | uniq -c | sort -rn | awk '{k=$5$6$7; if (c[k]++ < 3) print}' | cut -c1-300
