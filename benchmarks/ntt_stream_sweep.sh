# sweep of the batched-NTT grouping knobs (ntt.cu): groups in flight x bytes per group
for cfg in ${SWEEP:-"1 16" "2 16" "3 16" "2 32" "3 32" "4 32"}; do set -- $cfg; echo "== streams $1 group_mb $2"; STARK_NTT_STREAMS=$1 STARK_NTT_GROUP_MB=$2 python benchmarks/ntt_micro.py --logs 20,22 --batch 16 --lde 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['op'], d['log_n'], round(d['us'],1), round(d['hbm_frac_of_measured'],3))
"; done
