for L in 2 3 4; do for F in 32 64 128; do
echo -n "lanes $L first $F: "; CERB_PIPE_LANES=$L CERB_PIPE_FIRST=$F python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']))"
done; done
