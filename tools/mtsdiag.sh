for g in 64 128; do
  for lib in "" gpurun_variants/lib_base.so; do
    echo "== grid $g lib ${lib:-new}"
    CPFFT_B200_LIB=${lib:+$PWD/$lib} timeout 120 python bench.py --variant mts --grid $g --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --stress-leg-steps 0 --no-parity 2>&1 | tail -1 | python -c "
import sys, json
t=sys.stdin.read().strip()
try:
    l=json.loads(t); print('value %.4g'%l['value'], l['config'].get('newton_iters_per_step'), l['config'].get('G_K_dF_applies'), l['config'].get('mm10_local_failures'))
except Exception as e: print(t[-300:])
"
  done
done
