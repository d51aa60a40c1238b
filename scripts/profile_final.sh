#!/bin/bash
# The short evidence pass at the end of a session (under gpurun): headline bench, reference arm, launch list of the bench
# command, and the network configs with step graphs on and off.  Usage: bash scripts/profile_final.sh r1g
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference_n1.json 2>> $OUT/${TAG}_bench_n1.err
ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches_bench_steps2_warmup3.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
: > $OUT/${TAG}_networks_n1.jsonl
for c in 1 3 4 5; do
	for g in 0 1; do
		CATTL3_NO_GRAPH=$g timeout 200 python scripts/bench_networks.py --config $c 2>/dev/null | tail -1 |
			sed "s/\"impl\": \"b200\"/\"impl\": \"b200\", \"step_graphs\": $((1 - g))/" >> $OUT/${TAG}_networks_n1.jsonl
	done
done
for c in 1 3; do
	timeout 200 python scripts/bench_networks.py --config $c --impl reference --steps 2 --epochs 1 2>/dev/null | tail -1 >> $OUT/${TAG}_networks_n1.jsonl
done
timeout 200 python scripts/bench_networks.py --config 5 --impl reference --batch 8 --steps 1 --epochs 1 2>/dev/null | tail -1 >> $OUT/${TAG}_networks_n1.jsonl
cat $OUT/${TAG}_bench_n1.json | cut -c1-900
cut -c1-170 $OUT/${TAG}_networks_n1.jsonl
