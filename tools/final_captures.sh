#!/bin/bash
# Final evidence of a round, run on the GPU box under gpurun. Every stage writes its results under gpurun_out/ as soon as it
# is done (a call cut off by the GPU budget still brings the earlier stages back), most important first:
#   1. the GPU test suite and smoke();
#   2. one `ncu --set full` launch of each c2 bucket kernel -> summaries + profiles/traffic.json entries (keyed by the library's
#      source digest), each capture only after the same command exited 0 without ncu;
#   3. the same for the c4b (north-star config) and c3 mesh kernels;
#   4. the launch list of the headline bench command;
#   5. the default bench line (reads the fresh traffic.json).
# The .ncu-rep files stay in /tmp on the box: together they exceed what gpurun copies back.
set -u
TAG=${TAG:-r02z}
OUT=gpurun_out
T=/tmp/ncu_reports; mkdir -p $T $OUT
B="--no-extra --no-e2e --no-cpu-baseline --no-spot-check"
T0=$(date +%s)
stamp() { echo "$(( $(date +%s) - T0 )) s: $*" | tee -a $OUT/${TAG}_timing.txt; }
extra='import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]
keys=["dram__bytes_read.sum.per_second","dram__bytes_write.sum.per_second","l1tex__m_xbar2l1tex_read_bytes.sum","l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum","lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_hit.sum","lts__t_sectors_srcunit_tex_op_read_evict_last_lookup_miss.sum","lts__t_sector_op_read_hit_rate.pct","lts__t_sector_op_write_hit_rate.pct","sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    for k in keys:
        if k in h: print("   %-90s %18s %s" % (k, r[h.index(k)], u[h.index(k)]))'
cap() {  # name, kernel regex, launch-skip, count, bench args...
  local name=$1 pat=$2 skip=$3 cnt=$4; shift 4
  python bench.py "$@" $B > $OUT/${TAG}_${name}_plain.json 2>&1 && \
    ncu --set full --import-source on --clock-control none -k regex:$pat --launch-skip $skip -c $cnt -f -o $T/$name python bench.py "$@" $B > $OUT/${TAG}_${name}_ncu.log 2>&1
  stamp "$name capture rc=$?"
}

rm -f $OUT/${TAG}_timing.txt
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > $OUT/${TAG}_gputests.txt; stamp "gpu tests: $(tail -1 $OUT/${TAG}_gputests.txt)"
python __graft_entry__.py smoke 2>&1 | tail -1 > $OUT/${TAG}_smoke.txt; stamp "smoke: $(cat $OUT/${TAG}_smoke.txt)"

rm -f profiles/traffic.json
cap c2 bucket_ 6 2 --workload c2 --queries 4194304 --steps 1 --warmup 3
(python tools/ncusum.py $T/c2.ncu-rep; python tools/ncusrc.py $T/c2.ncu-rep bucket_eval 25; python tools/ncusrc.py $T/c2.ncu-rep bucket_scatter 25) > $OUT/${TAG}_c2_bucket_kernels_ncu_full.txt 2>&1
python tools/traffic_from_ncu.py $T/c2.ncu-rep bucket_eval c2 bucket_eval 4194304 "ncu --set full, one launch over a 4,194,304-query chunk"
python tools/traffic_from_ncu.py $T/c2.ncu-rep bucket_scatter c2_scatter bucket_scatter 4194304 "ncu --set full, one launch over a 4,194,304-query chunk"
cp profiles/traffic.json $OUT/${TAG}_traffic.json; stamp "c2 summarised"

cap c4b mesh_tensor_mma 20 1 --workload c4b --steps 1 --warmup 1
(python tools/ncusum.py $T/c4b.ncu-rep; ncu -i $T/c4b.ncu-rep --page raw --csv 2>/dev/null | python -c "$extra"; python tools/ncusrc.py $T/c4b.ncu-rep mesh_tensor_mma 25) > $OUT/${TAG}_c4b_mesh_tensor_mma_ncu_full.txt 2>&1
Q4=$(python -c "
import json; d=json.loads([l for l in open('$OUT/${TAG}_c4b_plain.json') if l.startswith('{')][-1]); print(d['roofline']['queries_per_launch'])")
python tools/traffic_from_ncu.py $T/c4b.ncu-rep mesh_tensor_mma c4b mesh_tensor_mma $Q4 "ncu --set full, one ring run of the 262,144-query step"
cp profiles/traffic.json $OUT/${TAG}_traffic.json; stamp "c4b summarised"

cap c3 mesh_tensor_mma 3 1 --workload c3 --steps 1 --warmup 3
(python tools/ncusum.py $T/c3.ncu-rep; ncu -i $T/c3.ncu-rep --page raw --csv 2>/dev/null | python -c "$extra"; python tools/ncusrc.py $T/c3.ncu-rep mesh_tensor_mma 25) > $OUT/${TAG}_c3_mesh_tensor_mma_ncu_full.txt 2>&1
python tools/traffic_from_ncu.py $T/c3.ncu-rep mesh_tensor_mma c3 mesh_tensor_mma 262144 "ncu --set full, one launch = one 262,144-query step"
cp profiles/traffic.json $OUT/${TAG}_traffic.json; stamp "c3 summarised"

python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_plain_c2.json 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_c2_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_launch_c2.log 2>&1
stamp "launch list rc=$?"
python tools/launchsum.py $OUT/${TAG}_c2_launches.csv > $OUT/${TAG}_c2_launches_summary.txt 2>&1

python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; stamp "default bench rc=$?"
ls -la $OUT | tail -30; du -sh $OUT
