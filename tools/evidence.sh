#!/bin/bash
# One-box evidence snapshot for profiles/: GPU tests, smoke, every bench line, the reference arm, the ncu launch list and
# the --set full capture of one forward.  Run on a B200 box from the repo root:  bash tools/evidence.sh r02w
# Writes gpurun_out/<tag>_*; copy what is to be judged into profiles/ (tools/ncu_summary.py makes the .md files).
tag=${1:-snap}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_gputest.log 2>&1; echo "rc=$?" >> $out/${tag}_gputest.log
python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1
python bench.py > $out/${tag}_bench_4096.json 2> $out/${tag}_bench.err
for c in c1 c3 c4 c5; do python bench.py --config $c > $out/${tag}_bench_$c.json 2>> $out/${tag}_bench.err; done
python bench.py --impl reference --steps 3 --warmup 3 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --videos 512 --steps 1 --no-cpu-baseline --no-modes --no-other-configs > $out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -o $out/${tag}_fwd -f \
    python tools/profile_forward.py 512 fp16x2 > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_gputest.log; tail -1 $out/${tag}_smoke.log; cut -c1-220 $out/${tag}_bench_4096.json
