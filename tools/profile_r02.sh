#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun; outputs land in gpurun_out/):
# one `ncu --set full` capture per kernel (after the plain run of the same command exited 0), summarised on
# the box (key metrics + per-source-line shares; the reports themselves are too big to bring back, only the
# one of the bench kernel is kept), the launch list of a short bench run, and the plain timings of the
# measurement tools.
set -u
out=gpurun_out
lib=sac-gat-her_transportationrl_b200/libb200assign.so
cap() {  # cap <name> <kernel regex> <skip> <keep report: 0|1> <command...>
  local name=$1 regex=$2 skip=$3 keep=$4; shift 4
  if [ -n "${ONLY:-}" ] && [[ ! " $ONLY " =~ " $name " ]]; then return; fi   # ONLY="assign_static gat_fwd" bash tools/profile_r02.sh
  "$@" > $out/r02_${name}_plain.log 2>&1 || { echo "plain run of $name failed"; return; }
  ncu --set full --import-source on --clock-control none -k regex:$regex -s $skip -c 1 -f -o $out/r02_$name "$@" > $out/r02_${name}_ncu.log 2>&1
  { echo "# $name: $*"; grep -h "^{" $out/r02_${name}_plain.log | tail -3; python tools/ncu_rawsum.py $out/r02_$name.ncu-rep;
    echo; python tools/ncu_lines.py $out/r02_$name.ncu-rep $lib $regex 1.5; } > $out/r02_${name}_ncu_summary.txt 2>&1
  [ "$keep" = 1 ] || rm -f $out/r02_$name.ncu-rep
  rm -f $out/r02_${name}_ncu.log
}
cap assign_static assign_static 3 1 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline
cap assign_fused_an assign_fused 1 0 python tools/profile_general.py an 1024 8
cap assign_fused_cs assign_fused 1 0 python tools/profile_general.py cs 148 4
cap features features_small 1 0 python tools/measure_features.py sf
cap features_an features_kernel 1 0 python tools/measure_features.py an
cap bpr bpr_kernel 2 0 python tools/measure_stages.py 1048576 16384
cap sssp sssp_kernel 2 0 python tools/measure_stages.py 65536 16384
cap aon aon_kernel 2 0 python tools/measure_stages.py 65536 16384
cap update update_kernel 2 0 python tools/measure_stages.py 65536 16384
cap gat_fwd gat_fwd 2 0 python tools/measure_gat.py 4096 4 256
cap gat_bwd gat_bwd 2 0 python tools/measure_gat.py 4096 4 256
[ -n "${ONLY:-}" ] && { du -sh $out; exit 0; }
python tools/measure_stages.py > $out/r02_stages.log 2>&1
python tools/measure_features.py sf an cs > $out/r02_features.log 2>&1
for h in 4 1; do python tools/measure_gat.py 8192 $h 256; done > $out/r02_gat.log 2>&1
python tools/measure_greedy.py > $out/r02_greedy.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/r02_launches_run.log 2>&1
du -sh $out
