cd /root/repo
( nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 250 > gpurun_out/r02_v17_clk.txt ) &
SMI=$!
sleep 1
timeout 100 tools/microbench/bin/oz_fwd_bench 3600 1 20 6 1
kill $SMI
echo "--- clocks (sm MHz, W, reasons) during the long run:"; sort gpurun_out/r02_v17_clk.txt | uniq -c | sort -rn | head -8
echo "--- spin backoff 40 ns:"; timeout 100 tools/microbench/bin/oz_fwd_bench_sleep 3600 1 20 6 | grep -v mismatch | cut -c1-100
