for i in 1 2; do for lib in "" examples/bench_scripts/libold.so; do PIME_B200_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-aux --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('wt lib=$lib', '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
"; done; done; for i in 1 2; do for lib in "" examples/bench_scripts/libold.so; do PIME_B200_LIB=$lib timeout 300 python bench.py --workload ph --steps 5 --warmup 3 --no-cpu-baseline --no-aux --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ph lib=$lib', '%.4g'%d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])
"; done; done
