#!/bin/bash
# One parameterised GPU session (replaces the per-run scratch scripts of round 1):
#   gpurun --timeout T -- 'bash scripts/gpu_session.sh <tag> <stage> [<stage> ...]'
# Every stage writes gpurun_out/<tag>_<stage>.log and echoes its tail; a failing stage does not stop the session.
# Stages: tests tests_x tests_new tests_pack tests_attn attn_quick smoke bench bench_fast bench_c3 scan sweep retrieve retrieve_ab
#         (ABSET="15=1 15=3 ...": same-box A/B through DRAG_DEBUG_SET) retrieve_b32 ref attn gemm epilogue launches ncu_hot ncu_attn ncu_stem ncu_vit
# Environment: NGPU (torchrun ranks), STEPS, TAIL (lines echoed per stage)
tag=$1; shift
mkdir -p gpurun_out
run() { local t=$1 name=$2; shift 2; local t0=$SECONDS; timeout $t "$@" > gpurun_out/${tag}_$name.log 2>&1; local rc=$?
        echo "rc=$rc secs=$((SECONDS-t0))" >> gpurun_out/${tag}_$name.log; echo "== $name"; tail -n ${TAIL:-6} gpurun_out/${tag}_$name.log | cut -c1-3000; return $rc; }
N=${NGPU:-1}
launch() { if [ "$N" -gt 1 ]; then echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; else echo "python"; fi; }
for stage in "$@"; do
  case $stage in
    tests)        run 1500 tests python -m pytest tests -m gpu -q --durations=10 -s ;;
    tests_x)      run 1500 tests python -m pytest tests -m gpu -x -q --durations=10 ;;
    tests_new)    run 900 tests_new python -m pytest tests/test_full_depth_gpu.py tests/test_cli_gpu.py tests/test_pipelines_gpu.py tests/test_retrieval_cli_gpu.py tests/test_topk_gpu.py -m gpu -q -s --durations=10 ;;
    tests_pack)   run 600 tests_pack python -m pytest tests/test_pipelines_gpu.py tests/test_sharded_nccl_gpu.py tests/test_topk_gpu.py -m gpu -q --durations=5 ;;
    tests_attn)   run 600 tests_attn python -m pytest tests/test_flux_gpu.py tests/test_stem_gpu.py tests/test_vit_gpu.py -m gpu -q -s --durations=5 ;;
    attn_quick)   run 150 attn_tests python -m pytest tests/test_flux_gpu.py -k "attention" -m gpu -q -x
                  run 120 attn python scripts/bench_attn.py ;;
    epilogue)     run 200 epilogue python scripts/bench_epilogue.py ;;
    smoke)        run 300 smoke python -c "import __graft_entry__ as g; g.smoke()" ;;
    bench)        run 1200 bench $(launch) bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 ;;
    bench_fast)   run 900 bench_fast $(launch) bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 --no-secondary --no-gpu-baseline ;;
    bench_c3)     run 600 bench_c3 $(launch) bench.py --gpus $N --workload c3 ;;
    scan)         run 300 scan $(launch) bench.py --gpus $N --workload scan ;;
    sweep)        run 600 sweep $(launch) bench.py --gpus $N --workload scan --sweep --steps 10 ;;
    retrieve)     run 400 retrieve $(launch) bench.py --gpus $N --workload retrieve ;;
    retrieve_ab)  for kv in ${ABSET:-"15=1" "15=2" "15=1" "15=2"}; do DRAG_DEBUG_SET=$kv run 400 retrieve_ab_${kv}_$SECONDS $(launch) bench.py --gpus $N --workload retrieve; done ;;
    retrieve_b32) run 400 retrieve_b32 $(launch) bench.py --gpus $N --workload retrieve --clip-model ViT-B/32 ;;
    ref)          run 400 ref python bench.py --impl reference --steps 1 --warmup 0 ;;
    attn)         run 300 attn python scripts/bench_attn.py ;;
    gemm)         run 300 gemm python scripts/bench_raster2.py ;;
    launches)     DRAG_BENCH_LAUNCH_LIST_ONLY=1 run 900 launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-secondary --no-gpu-baseline ;;
    ncu_hot)      run 600 ncu_hot ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|attention_tcgen05" -s 4 -c 4 -o gpurun_out/${tag}_hot python scripts/prof_kernels.py 4 ;;
    ncu_attn)     run 400 ncu_attn ncu --set full --clock-control none --import-source on -k regex:"attention_tcgen05" -s 2 -c 2 -o gpurun_out/${tag}_attn python scripts/prof_kernels.py 4 ;;
    ncu_stem)     run 300 ncu_stem ncu --set full --clock-control none --import-source on -k regex:"stem_stats" -s 1 -c 1 -o gpurun_out/${tag}_stem python scripts/prof_stem.py ;;
    ncu_vit)      run 400 ncu_vit ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gemm_bf16|attention_tcgen05|layernorm" -s 14 -c 14 --csv --log-file gpurun_out/${tag}_vit_plain.csv python scripts/prof_vit.py plain
                  run 400 ncu_vit_fold ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gemm_bf16|attention_tcgen05|layernorm" -s 10 -c 10 --csv --log-file gpurun_out/${tag}_vit_fold.csv python scripts/prof_vit.py fold ;;
    *)            echo "unknown stage $stage" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -2
exit 0
