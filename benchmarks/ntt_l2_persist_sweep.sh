for p in 0 32 64 96; do echo "== L2_PERSIST $p"; STARK_NTT_L2_PERSIST=$p python benchmarks/ntt_micro.py --logs 20,22 --batch 16 --lde 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['op'], d['log_n'], round(d['us'],1), round(d['hbm_frac_of_measured'],3))
"; done
echo "== L2_PERSIST 64 streams 3"; STARK_NTT_STREAMS=3 STARK_NTT_L2_PERSIST=64 python benchmarks/ntt_micro.py --logs 22 --batch 16 2>&1 | cut -c1-120
