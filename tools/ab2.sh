run() { # name env libpath
  env $2 VSG_LIB_PATH=$3 python bench.py --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), {k: round(v,3) for k,v in d['stages_ms_per_step'].items()})"
}
run fused "X=1" ""
run unfused "VSG_FUSE_FAST_BLUR=0" ""
run unfused_minb7 "VSG_FUSE_FAST_BLUR=0" $PWD/gpurun_variants/libvsg_minb7.so
