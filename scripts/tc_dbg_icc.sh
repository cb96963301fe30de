for d in 0 4 8 12 15; do
  echo "dbg=$d"
  UHSDR_B200_TC_DEBUG=$d timeout 120 ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gpu__time_duration.sum --clock-control none -k regex:rx_ssb_tc -c 1 python bench.py --channels 4096 --blocks 400 --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 0 --parity-channels 0 2>&1 | grep -E "icc_request|gpu__time" 
done
