run() { echo "== $1"; env $1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],2), 'clips/s', round(d['ms_per_step'],1), 'ms/step')"; }
run "SFB_X=1"
run "SFB_NO_EPI12=1"
run "SFB_LIB=$PWD/syncfusion_b200/libsfb_old_ff02634.so"
