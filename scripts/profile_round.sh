#!/bin/bash
# One GPU box, one pass: the evidence files of a round (see profiles/README.md).  Usage (under gpurun):
#   bash scripts/profile_round.sh r1e
# Writes gpurun_out/<tag>_*.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
timeout 120 ./scripts/tf32_peak > $OUT/${TAG}_tf32_peak.json 2>&1
timeout 60 ./scripts/tf32_peak 0.3 ts 64 >> $OUT/${TAG}_tf32_peak.json 2>&1
python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_n1.json 2>> $OUT/${TAG}_bench_n1.err
python scripts/bench_ops.py --double > $OUT/${TAG}_ops_sweep.jsonl 2> $OUT/${TAG}_ops.err
./scripts/dfma_peak > $OUT/${TAG}_dfma_peak.json 2>&1
[ -x ./scripts/dmma_peak ] && ./scripts/dmma_peak >> $OUT/${TAG}_dfma_peak.json 2>&1
python scripts/bench_ops.py --only conv --double --path 3 2>/dev/null | grep float64 > $OUT/${TAG}_ops_sweep_dfma.jsonl
for c in 1 3 4 5; do python scripts/bench_networks.py --config $c 2>/dev/null | tail -1; done > $OUT/${TAG}_networks_n1.jsonl
for c in 1 3; do python scripts/bench_networks.py --config $c --impl reference --steps 2 --epochs 1 2>/dev/null | tail -1; done >> $OUT/${TAG}_networks_n1.jsonl
# launch list of the bench command (every launch with its device time)
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches_bench_steps2_warmup3.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-networks --no-double > /dev/null 2>&1
# full captures: the three tcgen05 GEMM kernels of one training step (skip the warm-up launches)
$NCU --set full --import-source on -k regex:"tc_gather_gemm_kernel|tc_wgrad_kernel|tc_rows_gemm_kernel" -s 9 -c 3 -o $OUT/${TAG}_tc \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-networks --no-double > /dev/null 2>&1
# the double kernels of config 2: FP64 tensor cores (DMMA: what AUTO runs), then the DFMA kernels (--path 3)
$NCU --set full --import-source on -k regex:"dmma2?_gather_gemm_kernel|dmma2?_wgrad_kernel" -s 30 -c 4 -o $OUT/${TAG}_dmma \
	python scripts/bench_ops.py --only conv --double --reps 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"fma_gather_gemm_kernel|fma_wgrad_kernel" -s 0 -c 4 -o $OUT/${TAG}_dfma \
	python scripts/bench_ops.py --only conv --double --reps 1 --path 3 > /dev/null 2>&1
# the fused forward epilogues
$NCU --set full --import-source on -k regex:"tc_gather_gemm_kernel|bn_apply_act_kernel" -s 20 -c 6 -o $OUT/${TAG}_fused \
	python scripts/bench_ops.py --only fused --reps 1 > /dev/null 2>&1
# the memory-bound layers: duration and DRAM traffic of every launch of the per-op sweep
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
	-k regex:"act_fwd_kernel|act_bwd_kernel|pool_fwd_kernel|pool_bwd|bn_stats_partial|bn_apply_kernel|bn_bwd_partial|bn_bwd_apply|glue_kernel|opt_step_kernel|softmax|dropout|slice_rows" \
	-c 200 --csv --log-file $OUT/${TAG}_mem_kernels.csv python scripts/bench_ops.py --only mem --reps 1 > /dev/null 2>&1
# launch list of one config-4 epoch (ResNet modules, 224 x 224, batch 64)
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $OUT/${TAG}_launches_net4.csv \
	python scripts/bench_networks.py --config 4 --steps 2 --epochs 1 > /dev/null 2>&1
gzip -f $OUT/${TAG}_launches_net4.csv
# gpurun brings back at most 64 MiB: keep the raw pages as CSV, drop the big reports
for r in tc dmma dfma fused; do
	ncu -i $OUT/${TAG}_$r.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_raw_$r.csv 2>/dev/null
done
ncu -i $OUT/${TAG}_tc.ncu-rep --page source --csv > $OUT/${TAG}_ncu_source_tc.csv 2>/dev/null
rm -f $OUT/${TAG}_dmma.ncu-rep $OUT/${TAG}_dfma.ncu-rep $OUT/${TAG}_fused.ncu-rep
gzip -f $OUT/${TAG}_ncu_source_tc.csv
ls -la $OUT | tail -20; du -sh $OUT
