#!/bin/bash
# Network-level evidence of a round on one GPU box (under gpurun): every network config through the C++ batch loop with
# step graphs on and off, the reference on the host cores beside it (bounded samples), and the launch list of one
# eager config-5 epoch.  Usage: bash scripts/profile_networks.sh r1g   -> gpurun_out/<tag>_*
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
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
CATTL3_NO_GRAPH=1 timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 6000 --csv \
	--log-file $OUT/${TAG}_launches_net5.csv python scripts/bench_networks.py --config 5 --steps 2 --epochs 1 > /dev/null 2>&1
gzip -f $OUT/${TAG}_launches_net5.csv
ls -la $OUT | tail -5
