#!/bin/bash
# On the GPU box: profile ONE kernel of a command with `ncu --set full`, keep only compact text artefacts.
# usage: tools/ncu_box.sh <out-name> <kernel-regex> <env-steps-per-launch> <command...>
# Writes gpurun_out/<name>.txt (tools/ncu_summary.py), gpurun_out/<name>.source.csv.gz (per-instruction page) and
# gpurun_out/<name>.raw.csv; the .ncu-rep itself (35 MB: it embeds the whole cubin) stays on the box.
name=$1; regex=$2; steps=$3; shift 3
mkdir -p gpurun_out
"$@" > gpurun_out/$name.plain.log 2>&1 || { echo "plain run failed: $name"; tail -5 gpurun_out/$name.plain.log; exit 1; }
ncu --set full --clock-control none --import-source off -k regex:$regex -s 2 -c 1 -f -o /tmp/$name "$@" > gpurun_out/$name.ncu.log 2>&1
if [ -f /tmp/$name.ncu-rep ]; then
  python tools/ncu_summary.py /tmp/$name.ncu-rep $steps > gpurun_out/$name.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/$name.source.csv.gz
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  cat gpurun_out/$name.plain.log
  head -40 gpurun_out/$name.txt
else
  echo "no report for $name"; tail -5 gpurun_out/$name.ncu.log
fi
