for v in "$@"; do
  WBC_B200_LIB=$PWD/build_variants/libwbc_$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['roofline_fk_jac']['achieved'], d['roofline_fk_jac']['ms'], round(d['value']/1e6,2))"
done
