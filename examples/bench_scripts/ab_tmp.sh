for i in 1 2 3; do for lib in "" examples/bench_scripts/libtmema.so; do PIME_B200_LIB=$lib timeout 300 python bench.py --workload ph --steps 8 --warmup 3 --no-cpu-baseline --no-aux --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ph lib=$lib', '%.4g'%d['value'], '%.3f'%d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])
"; done; done
for lib in "" examples/bench_scripts/libtmema.so; do PIME_B200_LIB=$lib timeout 300 python bench.py --net-dim 128 --steps 8 --warmup 3 --no-cpu-baseline --no-aux --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('wt128 lib=$lib', '%.4g'%d['value'], '%.3f'%d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])
"; done
