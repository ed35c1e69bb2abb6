#!/bin/bash
# ncu captures behind profiles/r02_*: run on a B200 box (gpurun), one `ncu --set full` report per kernel family into
# gpurun_out/.  Each case runs plain first (B200_PROFILING.md: only profile a command that has just exited 0).
set -u
out=gpurun_out
mkdir -p $out
cap() {  # case  kernel-regex  launches-to-skip
  python tools/profile_cases.py $1 > $out/plain_$1.log 2>&1 || { echo "plain run of $1 failed"; tail -3 $out/plain_$1.log; return; }
  ncu --set full --clock-control none -k regex:$2 -s $3 -c 1 -f -o $out/r02_$1 \
      python tools/profile_cases.py $1 > $out/ncu_$1.log 2>&1
  # the raw page travels back as CSV (gpurun_out is capped at 64 MiB); the report itself is dropped
  ncu -i $out/r02_$1.ncu-rep --page raw --csv > $out/r02_$1_rawpage.csv 2>/dev/null && rm -f $out/r02_$1.ncu-rep
  rm -f $out/ncu_$1.log
  tail -1 $out/plain_$1.log
}
want=${1:-all}
sel() { [ "$want" = all ] || [[ " $want " == *" $1 "* ]]; }
sel shoot && cap shoot shoot_queue_kernel 1
sel conte && cap conte shoot_queue_kernel 1
sel spatial && cap spatial shoot_pair_kernel 0
sel dense512 && cap dense512 chan_eig_kernel_big2 2
sel dense256 && cap dense256 '^chan_eig_kernel$' 2
sel dense128 && cap dense128 chan_eig_kernel_small 2
sel dense64 && cap dense64 chan_eig_warp_kernel 2
sel band512 && cap band512 chan_fd_banded_warp_kernel 2
sel band512_small && cap band512_small chan_fd_banded_warp_kernel 2
sel band64 && cap band64 chan_fd_banded_warp_kernel 2
sel neutral && cap neutral chan_neutral_warp_kernel 0
sel blasius && cap blasius blasius_kernel 0
sel launches || exit 0
# launch list of the default bench command (shares, not absolutes)
python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --points 1048576 > $out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary --points 1048576 > $out/ncu_bench.log 2>&1
ls -la $out/*.ncu-rep | wc -l
