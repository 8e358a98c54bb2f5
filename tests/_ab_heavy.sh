for lib in variants/libgsr_prev.so default variants/libgsr_prev.so default; do
  if [ "$lib" = default ]; then unset GSR_LIB_PATH; else export GSR_LIB_PATH=$PWD/$lib; fi
  python bench.py --value-only --steps 30 --scale-mult 3.0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
done
