#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/overlap_exp.log; : > $L
r() { env "$@" timeout 200 python tools/overlap_exp.py $ARGS >> $L 2>gpurun_out/overlap_exp.err || tail -5 gpurun_out/overlap_exp.err >> $L; }
ARGS="--streams 1 --mb 256" r A=1
ARGS="--streams 1 --mb 256" r NIB_TC_SERP=1
ARGS="--streams 1 --mb 192" r NIB_TC_SERP=1
ARGS="--streams 1 --mb 128" r NIB_TC_SERP=1
ARGS="--streams 1 --mb 128" r A=1
ARGS="--streams 2 --mb 256" r A=1
ARGS="--streams 2 --mb 256" r NIB_TC_PAIRS=37
ARGS="--streams 2 --mb 256" r NIB_TC_PAIRS=37 NIB_TC_NO_PDL=1
ARGS="--streams 2 --mb 128" r NIB_TC_PAIRS=37 NIB_TC_NO_PDL=1
ARGS="--streams 2 --mb 128" r NIB_TC_PAIRS=37 NIB_TC_NO_PDL=1 NIB_TC_SERP=1
ARGS="--streams 2 --mb 256" r NIB_TC_PAIRS=50 NIB_TC_NO_PDL=1
ARGS="--streams 3 --mb 128" r NIB_TC_PAIRS=25 NIB_TC_NO_PDL=1
ARGS="--streams 2 --mb 256" r NIB_TC_NO_PDL=1
cat $L
