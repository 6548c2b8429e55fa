for v in "$@"; do
  WBC_B200_LIB=$PWD/build_variants/libwbc_$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,2), 'M/s e2e', round(d['e2e']['value']/1e6,2), d['launch'], d['verified'])"
done
