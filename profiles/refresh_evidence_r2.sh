#!/usr/bin/env bash
# Round 2 evidence (run through gpurun; outputs land in gpurun_out/ and are copied into profiles/ by hand).
# Every ncu run follows a plain run of the same command that exited 0.
set -u
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-shim --skip-tc"
python bench.py --steps 20 --warmup 3 > $O/r2_bench_n1.json 2> $O/bench_n1.err || { tail -5 $O/bench_n1.err; exit 1; }
$BENCH > $O/bench_short.json 2> $O/bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_bench.csv $BENCH > $O/ncu_launch.log 2>&1
python profiles/launch_summary.py $O/r2_launches_bench.csv "$BENCH" > $O/r2_launches_summary.txt
# the dominant kernel of the timed region: one k_frame_seq launch = 20 frames
ncu --set full --clock-control none --import-source on -k "regex:k_frame_seq" -s 7 -c 1 -f -o $O/r2_frame_seq $BENCH > $O/ncu_frame.log 2>&1
python profiles/ncu_table.py $O/r2_frame_seq.ncu-rep > $O/r2_frame_seq.txt 2>&1
python profiles/ncu_kernels.py $O/r2_frame_seq.ncu-rep --json $O/r2_traffic.json --frames-per-launch 20 >> $O/r2_frame_seq.txt 2>&1
# closed loop: the fused step kernel
NAV_RUN_TRACE=1 python profiles/prof_closed_loop.py 400 > $O/r2_closed_loop.txt 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_loop_step" -s 20 -c 2 -f -o $O/r2_loop_step python profiles/prof_closed_loop.py 60 > $O/ncu_loop.log 2>&1
python profiles/ncu_table.py $O/r2_loop_step.ncu-rep > $O/r2_loop_step.txt 2>&1
# kd build: time per call, then every launch of one 1 M build with its DRAM bytes (per-level traffic)
python profiles/prof_kdbuild.py > $O/r2_kdbuild.txt 2>&1
ncu --set full --clock-control none -k "regex:k_kd_" -s 60 -c 30 -f -o $O/r2_kdbuild_levels python profiles/prof_kdbuild.py 1000000 > $O/ncu_kdb.log 2>&1
python profiles/ncu_kernels.py $O/r2_kdbuild_levels.ncu-rep > $O/r2_kdbuild_levels.txt 2>&1
python profiles/prof_frame.py > $O/r2_frame_us.txt 2>&1
NAV_SEQ_LAUNCHES=1 python profiles/prof_frame.py >> $O/r2_frame_us.txt 2>&1
python profiles/prof_stencil.py 1000 5 > $O/r2_stencil_1000_frames.txt 2>&1
python profiles/prof_nn.py 1024 4096 16384 65536 1000000 10000000 > $O/r2_nn_sizes.txt 2>&1
python profiles/prof_pcie.py > $O/r2_pcie.txt 2>&1
python profiles/prof_nn_frame.py > $O/r2_nn_frame.txt 2>&1
python profiles/sass_opcodes.py > $O/r2_sass_opcodes.txt 2>&1
rm -f $O/*.ncu-rep.tmp
ls -la $O | tail -30
