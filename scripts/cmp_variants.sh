for v in auto lat thr; do
  for wl in bunny_goicp_toml bunny_goicp_certified; do
    if [ $v = auto ]; then unset GOICP_BNB_VARIANT; else export GOICP_BNB_VARIANT=$v; fi
    python bench.py --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$wl', 'ms/step %.2f'%d['ms_per_step'], 'bnb %.2f ms'%(1e3*d['seconds_bnb_kernels_per_step']), 'icp %.2f'%(1e3*d['seconds_icp_per_step']), 'frac %.3f'%d['roofline']['frac'])"
  done
done
