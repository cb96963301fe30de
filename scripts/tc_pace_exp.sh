for d in 0 60 120 200 300; do
  echo -n "pace=$d: "
  UHSDR_B200_TC_PACE=$d timeout 100 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 0 --parity-channels 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done
