for v in cur lane corr cur lane corr; do
  WBC_B200_LIB=$PWD/build_variants/$v/libwbc_b200.so python bench.py --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']
print('$v', 'value %.2f e2e %.2f steps %d cand %s solved %.4f kbar %.2f' % (d['value']/1e6, e['value']/1e6, e['steps'], {k:round(v/1e6,1) for k,v in e['host_path_candidates_steps_per_s'].items()}, e['solved_fraction_last_tick'], d['mean_qp_iterations']))"
done
