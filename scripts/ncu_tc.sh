#!/bin/bash
# ncu --set full of the three tcgen05 kernels of one bench step, for each CATTL3_TC_PAIRS setting given (default "0 3").
# Usage (under gpurun): bash scripts/ncu_tc.sh <tag> [masks...]   -> gpurun_out/<tag>_p<mask>_{raw,source}.csv(.gz)
TAG=${1:-rX}; shift
MASKS=${@:-0 3}
mkdir -p gpurun_out
for m in $MASKS; do
	CATTL3_TC_PAIRS=$m ncu --set full --clock-control none --import-source on -k regex:"tc_gather_gemm_kernel|tc_wgrad_kernel|tc_rows_gemm_kernel" -s 9 -c 3 \
		-o gpurun_out/${TAG}_p$m timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-networks --no-double > /dev/null 2>&1
	ncu -i gpurun_out/${TAG}_p$m.ncu-rep --page raw --csv > gpurun_out/${TAG}_p${m}_raw.csv 2>/dev/null
	ncu -i gpurun_out/${TAG}_p$m.ncu-rep --page source --csv > gpurun_out/${TAG}_p${m}_source.csv 2>/dev/null
	gzip -f gpurun_out/${TAG}_p${m}_source.csv
	rm -f gpurun_out/${TAG}_p$m.ncu-rep
done
ls -la gpurun_out | grep ${TAG}
