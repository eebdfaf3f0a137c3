#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/x_$name.json 2>/dev/null; }
run base A=1
run cw16 MDGAT_ATTN_CW=1
run cvt0 MDGAT_ATTN_CVT=0
run cvt1 MDGAT_ATTN_CVT=1
run cvt2 MDGAT_ATTN_CVT=2
run cvt3 MDGAT_ATTN_CVT=3
run nopdl MDGAT_PDL=0
env timeout 200 python bench.py --cuda-graph --no-cpu-baseline --no-eager --no-latency > gpurun_out/x_graph.json 2>/dev/null
