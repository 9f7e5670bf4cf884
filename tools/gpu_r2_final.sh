#!/bin/bash
# Round-2 validation of a tree on one B200: all GPU parity tests, smoke(), the default bench line, the reference
# arm, the ncu launch list of one step and --set full captures of the kernels that carry the step.
# usage: tools/gpu_r2_final.sh <tag> [nocap]
cd "$(dirname "$0")/.."
TAG=${1:-r02_final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "^c1:|^c5:" gpurun_out/${TAG}_pytest_gpu.log | head -40
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/bench_digest.py gpurun_out/${TAG}_bench.json || tail -c 2000 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench_ref.json
[ "$2" = "nocap" ] && exit 0
# launch list of one 3600-orientation design (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/gpu_ncu_factor.py 3600 > gpurun_out/${TAG}_launches.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.summary.txt 2>&1
head -24 gpurun_out/${TAG}_launches.summary.txt
rm -f gpurun_out/${TAG}_launches.csv.gz; gzip -f gpurun_out/${TAG}_launches.csv
cap() { # name regex skip count
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$1 \
      python tools/gpu_ncu_factor.py 3600 > gpurun_out/$1.log 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page details > gpurun_out/$1.details.txt 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  echo "== $1"; grep -E "^  [a-z].*\(|Duration|Registers Per|Achieved Occ|Executed Ipc Active|No Eligible|Mem Busy|Max Bandwidth|DRAM Throughput" gpurun_out/$1.details.txt | head -40
  rm -f gpurun_out/$1.ncu-rep
}
cap ${TAG}_oz "ozaki_gemm_kernel" 40 2
cap ${TAG}_tsqr "tsqr_sep_kernel" 1 1
cap ${TAG}_svdclip "svdclip_kernel" 1 1
cap ${TAG}_bwd "bwd_fused_kernel|chain_bwd_sep_kernel" 3 2
cap ${TAG}_gram "gram_sweep_kernel" 1 1
python tools/ncu_traffic.py gpurun_out/${TAG}_oz.raw.csv gpurun_out/${TAG}_oz_traffic.json "ncu --set full --clock-control none, ozaki_gemm_kernel<6> launches 41-42 of a 3600-orientation design (forward EpiPhaseSliceFix (integer drain + integer functor, no FP64 instruction, 4-byte stores) 128x80 tiles with 20 epilogue warps, backward EpiStoreF64 128x80 tiles); tools/gpu_r2_final.sh ${TAG}" > /dev/null 2>&1
ls -la gpurun_out/ | grep ${TAG} | awk '{print $5, $9}'
