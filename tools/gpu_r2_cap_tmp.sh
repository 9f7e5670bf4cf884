cd /root/repo
out=gpurun_out/r02_v27_render16
AB_ROUTES=2 timeout 400 ncu --set full --clock-control none --import-source on -k "regex:fused_render16_kernel" -s 1 -c 1 -f -o $out python tools/gpu_render_ab.py 60 > $out.log 2>&1
ncu -i $out.ncu-rep --page details > $out.details.txt 2>/dev/null
ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
grep -E "Duration|Registers Per|Achieved Occ|Executed Ipc Active|No Eligible|Mem Busy|Max Bandwidth|DRAM Throughput|L2 Hit|L1/TEX Hit|Bank|bank|cycles being stalled|Local|Executed Instructions  " $out.details.txt | head -40
rm -f $out.ncu-rep
