#!/bin/bash
# ncu captures of the round (run under gpurun, ONE GPU).
#  1. one single-pass list of EVERY launch of a default bench.py run with its duration and DRAM
#     bytes (no replay: the kernels run once, at the bench's own sizes)  -> r02_launches_bench.csv
#  2. one `--set full` capture per dominant kernel, summarised on the box (raw + details pages);
#     the reports themselves are dropped except the ones named in KEEP (gpurun_out/ carries 64 MiB).
# tools/ncu_traffic.py turns the results into profiles/r02_dram_traffic.json and summaries.
cd "${GRAFT_REPO_ROOT:-.}"
OUT=gpurun_out
mkdir -p $OUT
N="--clock-control none"
KEEP="zonal_select_main smooth_fast"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M $N --csv --log-file $OUT/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --leg-steps 3 --e2e-steps 2 > $OUT/r02_launches_bench.json 2> $OUT/r02_launches_bench.err
full() {  # name, kernel regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --set full $N --import-source on -k regex:$regex -s $skip -c 1 -f -o $OUT/r02_$name "$@" > $OUT/r02_$name.log 2>&1
  if [ -f $OUT/r02_$name.ncu-rep ]; then
    ncu -i $OUT/r02_$name.ncu-rep --page raw --csv > $OUT/r02_${name}_raw.csv 2>/dev/null
    ncu -i $OUT/r02_$name.ncu-rep --page details > $OUT/r02_${name}_ncu_details.txt 2>/dev/null
    case " $KEEP " in *" $name "*) ;; *) rm -f $OUT/r02_$name.ncu-rep ;; esac
  fi
}
full eval_specialised gm_fused 3 python bench.py --steps 5 --warmup 3 --profile
full zonal_reduce_warp zonal_reduce_warp 2 python tools/bench_kernels.py --only zonal_mean --zonal-size 40000 --zonal-grid 316 --iters 3
full zonal_select_bracket zonal_select_bracket 3 python tools/bench_kernels.py --only zonal_p90 --zonal-size 40000 --zonal-grid 316 --iters 3
full zonal_select_main zonal_select_main 3 python tools/bench_kernels.py --only zonal_p90 --zonal-size 40000 --zonal-grid 316 --iters 3
full zonal_select_final zonal_select_final 3 python tools/bench_kernels.py --only zonal_p90 --zonal-size 40000 --zonal-grid 316 --iters 3
full rasterize_tile rasterize_tile 2 python tools/bench_kernels.py --only rasterize --iters 3
full smooth_fast smooth_fast 2 python tools/bench_kernels.py --only smooth --scale 2 --iters 3
full moving_max_block moving_max_block 2 python tools/bench_kernels.py --only movingmax_11 --scale 2 --iters 3
full hillshade_quad hillshade_quad 2 python tools/bench_kernels.py --only hillshade --scale 2 --iters 3
full temporal_stream temporal_stream 2 python tools/bench_kernels.py --only temporal_sum --temporal-frames 64 --temporal-size 8192 --temporal-stats sum --iters 4
full temporal_cumulative temporal_cumulative_stream 2 python bench.py --steps 5 --warmup 3 --legs chain,temporal --temporal-frames 64 --size 4096
full temporal_moments temporal_moments_stream 2 python bench.py --steps 5 --warmup 3 --legs chain,temporal --temporal-frames 64 --size 4096
full temporal_sort temporal_sort_reg 2 python bench.py --steps 5 --warmup 3 --legs chain,temporal --temporal-frames 64 --size 4096
ls -la $OUT | grep r02_ | tail -50
du -sh $OUT
