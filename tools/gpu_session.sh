#!/usr/bin/env bash
# One gpurun call that produces everything a round needs from the GPU, in the order that matters if the call is cut
# short: parity tests, smoke, the bench line, the CPU arm, the ncu launch list of the bench command, one ncu --set full
# capture of the headline kernel families, the microbenchmarks.  Everything lands under gpurun_out/ (merged back by
# gpurun); copy what should be judged into profiles/ afterwards.
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session.sh'            # full session (~25 min)
#   /usr/local/graft/bin/gpurun --timeout 900  -- 'bash tools/gpu_session.sh tests bench' # selected stages
#
# Stages: tests smoke bench ref launches ncu micro shares small   (default: all but "small")
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
STAGES="${*:-tests smoke bench ref launches ncu micro shares}"
has() { [[ " $STAGES " == *" $1 "* ]]; }
log() { echo "[gpu_session $(date +%H:%M:%S)] $*" | tee -a gpurun_out/session.log; }

if has tests; then
  log "pytest -m gpu"
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  log "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"
fi
if has smoke; then
  log "smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  log "smoke exit $? : $(tail -1 gpurun_out/smoke.log)"
fi
if has bench; then
  log "bench N=1"
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  log "bench exit $? : $(head -c 300 gpurun_out/bench_n1.json)"
  log "bench N=1, batches assembled on the device (opt-in e2e input)"
  timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --e2e-input device_shards \
      > gpurun_out/bench_n1_device_shards.json 2> gpurun_out/bench_n1_device_shards.err
  log "bench (device shards) exit $?"
fi
if has ref; then
  log "reference arm"
  timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  log "reference exit $?"
fi
if has launches; then
  # per-launch durations of the library's kernels over the same bench command (cold-cache, serialised: shares only)
  log "ncu launch list"
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'oodfq|fq_|bn_|res_tail|s2d|weight_fq|minmax|energy|crop_resize|gap_' \
      -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-baselines --no-other-configs --graph off > gpurun_out/launches_run.log 2>&1
  log "launch list exit $?"
  python tools/launch_list.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1 || true
fi
if has ncu; then
  # one full capture per headline family, on the microbenchmark (a single tensor per launch, a few launches)
  log "ncu --set full (tail, bn, stem, s2d, fq, single-pass calibration)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'res_tail|bn_nhwc|bn_pool|s2d_stem|fq_flat|act_calib_onchip|gap_' -c 100 \
      -o /tmp/full_kernels -f python tools/microbench.py --only fq,tail,bn_fwd,bn_bwd,pool,s2d,calib_stats --shapes 0,1 --iters 1 --warmup 0 --flush none \
      > gpurun_out/ncu_full.log 2>&1
  log "ncu full exit $?"
  # the .ncu-rep itself (sources included) exceeds what gpurun copies back: it stays on the box, the tables come home
  ncu -i /tmp/full_kernels.ncu-rep --page raw --csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
      > gpurun_out/full_kernels_raw.csv 2>/dev/null || true
  python tools/ncu_traffic.py gpurun_out/full_kernels_raw.csv "profiles/r2_full_kernels_raw.csv (ncu --set full on tools/microbench.py, this build)" \
      > gpurun_out/ncu_traffic.json 2> gpurun_out/ncu_traffic.err || true
fi
if has micro; then
  log "microbench (read-flush, CUDA events)"
  timeout 900 python tools/microbench.py --only copy,fq,calib,calib_stats,tail,pool,s2d,bn_fwd,bn_bwd,stats_nhwc,weights,augment --flush read \
      --json gpurun_out/microbench.json > gpurun_out/microbench.txt 2>&1
  log "microbench exit $?"
  # event timestamps tick every ~2 us on this GPU: the small planes (and the ResNet-20 shapes 5-7) are timed by CUPTI
  log "microbench (read-flush, CUPTI kernel durations)"
  timeout 900 python tools/microbench.py --only copy,fq,calib,calib_stats,tail,pool,s2d,bn_fwd,bn_bwd,stats_nhwc,weights --shapes 0,1,2,3,4,5,6,7 \
      --flush read --cupti --iters 15 --json gpurun_out/microbench_cupti.json > gpurun_out/microbench_cupti.txt 2>&1
  log "microbench (cupti) exit $?"
fi
if has shares; then
  # where every workload's step goes (CUPTI), and the library's launches of the headline step per tensor size
  for w in imagenet_resnet18_w4a4 cifar100_resnet20_w4a4 pathmnist_resnet18_w2a2 distill_imagenet_resnet18_w4a4; do
    log "step share $w"
    timeout 400 python tools/step_profile.py --workload $w --steps 2 --top 60 --out gpurun_out/step_share_$w.txt > /dev/null 2>&1
  done
  timeout 400 python tools/step_profile.py --by-shape --steps 3 --out gpurun_out/step_by_shape.txt > /dev/null 2>&1
  log "step shares exit $?"
fi
if has small; then
  for w in cifar100_resnet20_w4a4 pathmnist_resnet18_w2a2 distill_imagenet_resnet18_w4a4; do
    log "bench $w"
    timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  done
fi
log "done"
