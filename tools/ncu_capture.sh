#!/bin/bash
# Round-2 ncu captures (run under gpurun): one `--set full` capture of the serving kernel per
# workload / content class + the launch list of the headline command.  Summaries: tools/ncu_summary.py.
set -u
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, bench args...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s "$skip" -c 1 -f \
      -o "gpurun_out/r2_$name" python bench.py --steps 3 --warmup 3 --profile --batch 16 "$@" \
      > "gpurun_out/r2_$name.log" 2>&1 || echo "capture $name failed"
  # summarise on the box and drop the 15 MB report (gpurun_out/ is capped at 64 MiB)
  if [ -f "gpurun_out/r2_$name.ncu-rep" ]; then
    python tools/ncu_summary.py "gpurun_out/r2_$name.ncu-rep" 132710400 "gpurun_out/r2_ncu_$name.md" \
      && rm -f "gpurun_out/r2_$name.ncu-rep"
  fi
}
for c in grad noise rand; do
  cap colorlut65_4k_$c vf_map_tile 2 --workload colorlut65_4k --content $c --option lut.path=4
done
cap hsvfilter_4k_noise_table vf_map_tile 2 --workload hsvfilter_4k --content noise --hsv-path 2
cap hsvfilter_4k_rand_compute vf_map_vec 2 --workload hsvfilter_4k --content rand --hsv-path 1
cap hsvdetector_4k_noise_table vf_map_tile 2 --workload hsvdetector_4k --content noise --hsv-path 2
cap colorlut33_4k_rgba64_grad vf_map_vec 2 --workload colorlut33_4k_rgba64 --content grad --option lut.path=4
cap colorlut33_4k_rgba64_noise vf_map_vec 2 --workload colorlut33_4k_rgba64 --content noise --option lut.path=4
cap colorlut33_4k_rgba64_noise_direct vf_map_vec 2 --workload colorlut33_4k_rgba64 --content noise --option lut.path=1
# launch list of the headline command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_colorlut65_4k.csv python bench.py --steps 5 --warmup 3 --profile --option lut.path=4 \
    > gpurun_out/r2_launches.log 2>&1
ls -la gpurun_out/*.md
