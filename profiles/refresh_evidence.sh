#!/usr/bin/env bash
# Regenerates the round's evidence files on a GPU box (run through gpurun; outputs land in gpurun_out/
# and are copied into profiles/ by hand).  Every ncu run follows a plain run of the same command.
set -u
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 20 --warmup 3 --skip-cpu --skip-batched"
python bench.py --steps 200 --warmup 5 --cpu-frames 20 > $O/r1_bench_n1.json 2> $O/bench_n1.err || exit 1
$BENCH > $O/bench_short.json 2> $O/bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r1_launches_bench.csv $BENCH > $O/ncu_launch.log 2>&1
python profiles/launch_summary.py $O/r1_launches_bench.csv "$BENCH" > $O/r1_launches_summary.txt
ncu --set full --clock-control none --import-source on -k "regex:k_frame_match|k_frame_map" -s 30 -c 2 -f -o $O/frame $BENCH > $O/ncu_frame.log 2>&1
python profiles/ncu_table.py $O/frame.ncu-rep > $O/r1_r1_frame.txt 2>&1
python profiles/prof_nn.py 1000000 > $O/nn_1m.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k "regex:k_kd_nn" -s 4 -c 1 -f -o $O/kdnn python profiles/prof_nn.py 1000000 > $O/ncu_kdnn.log 2>&1
python profiles/ncu_table.py $O/kdnn.ncu-rep > $O/r1_r1_kdnn.txt 2>&1
python profiles/prof_nn_frame.py > $O/r1_nn_frame.txt 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_kd_nn" -s 4 -c 1 -f -o $O/kdnn_room python profiles/prof_nn_frame.py > $O/ncu_kdnn_room.log 2>&1
python profiles/ncu_table.py $O/kdnn_room.ncu-rep > $O/r1_r1_kdnn_room.txt 2>&1
python profiles/prof_nn.py 1024 4096 16384 65536 1000000 10000000 > $O/r1_nn_sizes.txt 2>&1
python profiles/prof_kdbuild.py > $O/r1_kdbuild.txt 2>&1
python profiles/prof_tc.py 131072 256 1024 4096 16384 65536 > $O/r1_tc_sizes.txt 2>&1
python profiles/prof_csv.py > $O/r1_csv_rows.txt 2>&1
python profiles/prof_stencil.py 1000 5 > $O/r1_stencil_1000_frames.txt 2>&1
python profiles/prof_stencil.py 256 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_labels_tma" -s 2 -c 1 -f -o $O/stencil python profiles/prof_stencil.py 256 2 > $O/ncu_stencil.log 2>&1
python profiles/ncu_table.py $O/stencil.ncu-rep > $O/r1_r1_stencil.txt 2>&1
python profiles/prof_pcie.py > $O/r1_pcie.txt 2>&1
rm -f $O/*.ncu-rep.tmp
ls -la $O
