for rep in 1 2; do for v in "X=1" "SFB_NO_PDL=1"; do
  echo "$v default: $(env $v python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"],2))')"
done; done
for v in "X=1" "SFB_NO_PDL=1"; do
  echo "$v cfg: $(env $v python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --scale 2.0 2>/dev/null | grep '^{' | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"],2))')"
  echo "$v long: $(env $v python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --length 524288 --batch 4 2>/dev/null | grep '^{' | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"],2))')"
  echo "$v eager: $(env $v SFB_GRAPH=0 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c 'import sys,json; print(round(json.loads(sys.stdin.read())["value"],2))')"
done
