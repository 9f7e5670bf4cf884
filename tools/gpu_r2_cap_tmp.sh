cd /root/repo
out=gpurun_out/r02_v20_bwd_fused
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:bwd_fused_kernel" -s 3 -c 1 -f -o $out python tools/gpu_ncu_factor.py 3600 > $out.log 2>&1
ncu -i $out.ncu-rep --page details > $out.details.txt 2>/dev/null
ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
grep -E "Duration|Registers Per|Achieved Occ|Theoretical Occ|Executed Ipc Active|No Eligible|Mem Busy|Max Bandwidth|DRAM Throughput|L2 Hit|L1/TEX Hit|Bank|cycles being stalled|Block Limit|Shared Memory Config|Dynamic Shared|Executed Instructions  " $out.details.txt | head -40
rm -f $out.ncu-rep
