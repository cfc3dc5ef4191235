#!/usr/bin/env bash
# Round-2 evidence run on one B200: launch list, --set full captures, filter sweep, bench lines of every workload.
#   gpurun --timeout 1500 -- 'bash tools/gpu_collect_r02.sh'
set -u
O=gpurun_out/r02
mkdir -p $O
X="--metrics lts__t_bytes.sum,l1tex__t_bytes.sum,lts__t_sectors.sum"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
for wl in room roomu; do
  timeout 500 ncu --profile-from-start off --set full $X --clock-control none --import-source on -k regex:'k_trace' -c 24 -f -o $O/trace_$wl python tools/prof_frame.py $wl > $O/trace_$wl.log 2>&1
done
timeout 500 ncu --profile-from-start off --set full $X --clock-control none --import-source on -k regex:'k_shade|k_tail|k_generate|k_accumulate|k_wave_end|k_reproject|k_denoise' -c 24 -f -o $O/other_room python tools/prof_frame.py room > $O/other_room.log 2>&1
# summaries are made on the box: gpurun brings back at most 64 MiB, the reports are 30-50 MB each (only the C2 trace report travels)
for wl in room roomu; do
  python tools/ncu_summary.py report $O/trace_$wl.ncu-rep > $O/trace_${wl}_full.md
  python tools/ncu_summary.py traffic $O/trace_$wl.ncu-rep $wl $O/traffic.json > /dev/null
done
python tools/ncu_summary.py report $O/other_room.ncu-rep > $O/shade_filters_full.md
rm -f $O/other_room.ncu-rep $O/trace_roomu.ncu-rep
python tools/bench_filters.py 20 > $O/filters_sweep.jsonl 2> $O/filters_sweep.err
for wl in room roomu city city400 bricks cornell; do
  timeout 400 python bench.py --workload $wl > $O/n1_${wl}_tiles.json 2> $O/n1_${wl}_tiles.err
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/n1_room_reference.json 2> $O/n1_room_reference.err
ls -la $O
