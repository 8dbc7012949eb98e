#!/usr/bin/env bash
# GPU session 20 of round 2 (one B200): the build with the out-of-band step skip — full GPU suite, smoke, ncu of the cloud kernels
set -u
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:render_rays -s 1 -c 1 -f"
for w in cfg3A cfg3C cfg4A cfg4C; do
  timeout 300 $NCU -o $O/prof_${w}_tiled python profiles/prof_one.py $w tiled > $O/ncu_$w.log 2>&1
done
for f in $O/prof_*.ncu-rep; do python profiles/ncu_summary.py $f > ${f%.ncu-rep}.summary.txt 2>&1; done
python profiles/extract_facts.py $O r02 > $O/roofline_traffic_r02.json 2> $O/extract_facts.err
rm -f $O/prof_*.ncu-rep
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/roofline_traffic_r02.json | head -60
