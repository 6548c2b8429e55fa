# A/B of library builds under build_variants/ (WBC_B200_LIB override): bash tools/ab.sh name1 name2 ...   ("main" = the in-tree library)
for v in "$@"; do
  if [ "$v" = main ]; then lib=""; else lib=$PWD/build_variants/libwbc_$v.so; fi
  WBC_B200_LIB=$lib python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_err_$v.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,2), 'M/s e2e', round(d['e2e']['value']/1e6,2), d['launch'], d['verified'])
except Exception as e:
    print('$v', 'FAILED', e)"
done
