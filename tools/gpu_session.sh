#!/bin/bash
# Standard GPU-box sequences, sized from the timings measured in round 1 (one B200, this image):
#   getting a box + pushing the repo ~20 s; no-torch python start ~3 s; first `import torch` ~30 s;
#   pytest -m gpu (whole suite) ~125 s; bench.py --steps 2 --warmup 3 --no-cpu-baseline ~60 s.
# Everything writes under gpurun_out/ (merged back by gpurun).  Usage, from the repo root:
#   gpurun --timeout 240 -- 'tools/gpu_session.sh tests'         whole GPU suite
#   gpurun --timeout 120 -- 'tools/gpu_session.sh quick'         parity + golden + bit-identity tests (~15 s)
#   gpurun --timeout 200 -- 'tools/gpu_session.sh bench TAG'     bench line -> gpurun_out/TAG_bench256.json
#   gpurun --timeout 200 -- 'tools/gpu_session.sh launches TAG'  ncu launch list of one bench step
#   gpurun --timeout 120 -- 'tools/gpu_session.sh ncu-cg TAG'    ncu --set full of the kernels of one CG iteration
#   gpurun --timeout 300 -- 'tools/gpu_session.sh ncu-update TAG' ncu --set full of k_update_mm10 / k_pk1_tangent
#   gpurun --timeout 120 -- 'tools/gpu_session.sh ab'            A/B of the kernel-variant switches (tools/ab_iz.py)
#   gpurun --timeout 150 -- 'tools/gpu_session.sh ab-lf'         A/B of the lattice-frame residual variant of k_update_mm10 (CPFFT_MM10_LF=1)
#   gpurun --timeout 600 -- 'tools/gpu_session.sh first TAG'    first call of a round: new GPU tests, the three unmeasured variants, bench
set -u
MODE=${1:-quick}; TAG=${2:-rXX}
mkdir -p gpurun_out
case "$MODE" in
  tests)
    timeout 220 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee gpurun_out/${TAG}_gpu_tests.log ;;
  quick)
    timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_golden_decks.py \
        "tests/test_gpu_spectral.py::test_kernel_variants_are_bit_identical" -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_quick.log ;;
  bench)
    timeout 180 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench256.json 2> gpurun_out/${TAG}_bench256.err
    tail -c 400 gpurun_out/${TAG}_bench256.json ;;
  launches)
    timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_${TAG}.csv \
        python bench.py --grid 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --stress-leg-steps 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1
    python tools/ncu_summary.py gpurun_out/launches_${TAG}.csv | head -24 ;;
  ncu-cg)
    timeout 100 ncu --set full --clock-control none --import-source on -k 'regex:^(k_fz|k_fyf|k_fx|k_fyi|k_iz_pipe|k_cg_update_r)$' \
        -s 16 -c 6 -f -o gpurun_out/prof_${TAG}_cg python tools/ncu_cg.py 256 > gpurun_out/${TAG}_ncu_cg.log 2>&1
    tail -3 gpurun_out/${TAG}_ncu_cg.log; ls -la gpurun_out/prof_${TAG}_cg.ncu-rep ;;
  ncu-update)
    timeout 280 ncu --set full --clock-control none --import-source on -k 'regex:^(k_update_mm10[a-z_]*|k_pk1_tangent)$' -s 14 -c 2 -f \
        -o gpurun_out/prof_${TAG}_update python bench.py --grid 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --stress-leg-steps 0 > gpurun_out/${TAG}_ncu_update.log 2>&1
    tail -3 gpurun_out/${TAG}_ncu_update.log; ls -la gpurun_out/prof_${TAG}_update.ncu-rep ;;
  ab)
    timeout 100 python tools/ab_iz.py 256 10 2>&1 | tail -8 | tee gpurun_out/${TAG}_ab256.log ;;
  ab-lf)    # k_update_mm10 vs k_update_mm10_lf (residual slip loop in the lattice frame): time, checksum, local iterations
    (timeout 60 python tools/time_update.py 128; CPFFT_MM10_LF=1 timeout 60 python tools/time_update.py 128) 2>&1 | tail -4 | tee gpurun_out/${TAG}_ab_lf.log ;;
  mm10ab)   # A/B of libcpfft_b200.so variants built by tools/build_variants.py (gpurun_variants/lib_*.so): material sweep at 128^3
    for lib in gpurun_variants/lib_*.so; do CPFFT_B200_LIB=$PWD/$lib timeout 90 python tools/time_update.py 128 2>&1 | tail -1; done | tee gpurun_out/${TAG}_mm10ab.log ;;
  mm10ab2)  # round-2 k_update_mm10 changes: HEAD~ library vs the new one with its switches, then the LF-unroll variants
    L=$PWD/gpurun_variants
    ( CPFFT_B200_LIB=$L/lib_base.so timeout 90 python tools/time_update.py 128 2>&1 | tail -1
      CPFFT_B200_LIB=$L/lib_new.so CPFFT_MM10_UNI=0 timeout 90 python tools/time_update.py 128 2>&1 | tail -1
      CPFFT_B200_LIB=$L/lib_new.so timeout 90 python tools/time_update.py 128 2>&1 | tail -1
      CPFFT_B200_LIB=$L/lib_new.so CPFFT_MM10_LF=1 timeout 90 python tools/time_update.py 128 2>&1 | tail -1
      for v in ${3:-lfu2 lfu3 lfu4}; do CPFFT_B200_LIB=$L/lib_$v.so CPFFT_MM10_LF=1 timeout 90 python tools/time_update.py 128 2>&1 | tail -1; done
    ) | tee gpurun_out/${TAG}_mm10ab.log ;;
  mm10ab3)  # every library variant under gpurun_variants/, default and lattice-frame residual
    for lib in gpurun_variants/lib_*.so; do
      CPFFT_B200_LIB=$PWD/$lib timeout 90 python tools/time_update.py ${3:-128} 2>&1 | tail -1
      CPFFT_B200_LIB=$PWD/$lib CPFFT_MM10_LF=1 timeout 90 python tools/time_update.py ${3:-128} 2>&1 | tail -1
    done | tee gpurun_out/${TAG}_mm10ab.log ;;
  mm10ab4)  # variants x (fcc, bcc48 with and without the lattice-frame Jacobian); FP64 latency microbenchmark
    ( [ -x gpurun_variants/fp64_latency ] && timeout 60 gpurun_variants/fp64_latency
      for lib in gpurun_variants/lib_*.so; do
        CPFFT_B200_LIB=$PWD/$lib timeout 90 python tools/time_update.py 128 2>&1 | tail -1
        CPFFT_B200_LIB=$PWD/$lib timeout 90 python tools/time_update.py 128 bcc48 2>&1 | tail -1
      done
    ) | tee gpurun_out/${TAG}_mm10ab.log ;;
  izab)     # A/B of the inverse z pass occupancy variants (gpurun_variants/lib_iz*.so) at 256^3
    for lib in gpurun_variants/lib_iz*.so; do CPFFT_B200_LIB=$PWD/$lib timeout 90 python tools/time_apply.py 256 2>&1 | tail -2; done | tee gpurun_out/${TAG}_izab.log ;;
  scale)    # N-GPU call (gpurun --gpus N): correctness against the 1-GPU run, then bench.py as the driver launches it
    NG=${3:-8}; K=${4:-3}; W=${5:-3}
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
    timeout 150 $TR tools/multi_gpu_check.py --grid 64 --steps 3 2>&1 | grep -E "^\{|rror" | tee gpurun_out/${TAG}_check_${NG}gpu.log
    timeout 900 $TR bench.py --gpus $NG --steps $K --warmup $W > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err
    tail -3 gpurun_out/${TAG}_bench_${NG}gpu.err
    python - <<PY
import json
l=json.loads(open("gpurun_out/${TAG}_bench_${NG}gpu.json").read().strip().splitlines()[-1])
print("N", l["n_gpus"], "grid", l["config"]["grid"], "value %.4g" % l["value"], "e2e %.4g" % l["e2e"]["value"], "stress leg %.4g" % (l["stress_bc_leg"] or {}).get("value", 0),
      "parity", (l["parity"] or {}).get("ok"), {k: round(v["ms_per_launch"], 3) for k, v in l["stages"].items()})
PY
    ;;
  mgpu)     # N-GPU call (gpurun --gpus N): correctness against the 1-GPU run, then a short strain-BC bench
    NG=${3:-2}; GRID=${4:-320}
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
    (timeout 120 $TR tools/multi_gpu_check.py --grid 64 --steps 3) 2>&1 | grep -E "^\{|rror" | tee gpurun_out/${TAG}_mgpu${NG}_check.log
    for ch in 1; do ct=store
      timeout 300 $TR bench.py --gpus $NG --grid $GRID --steps 3 --warmup 3 --stress-leg-steps 0 --no-parity \
          > gpurun_out/${TAG}_bench${GRID}_${NG}gpu_chunks${ch}_${ct}.json 2> gpurun_out/${TAG}_bench${GRID}_${NG}gpu_chunks${ch}_${ct}.err
      python - <<PY
import json
l=json.loads(open("gpurun_out/${TAG}_bench${GRID}_${NG}gpu_chunks${ch}_${ct}.json").read().strip().splitlines()[-1])
print("xfer", "$ct", "chunks", $ch, "value %.4g" % l["value"], "e2e %.4g" % l["e2e"]["value"], {k: round(v["ms_per_launch"], 3) for k, v in l["stages"].items()})
PY
    done | tee gpurun_out/${TAG}_mgpu${NG}_pipeline.log ;;
  first)    # everything written without a GPU, cheapest first; every leg has its own timeout and log
    timeout 200 python -m pytest tests/test_zz_gpu_new_features.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_zz_tests.log
    "$0" ab-lf "$TAG"
    timeout 150 python bench.py --variant mts --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench256_mts.json 2> gpurun_out/${TAG}_bench256_mts.err
    tail -c 300 gpurun_out/${TAG}_bench256_mts.json
    "$0" bench "$TAG" ;;
  *) echo "unknown mode $MODE"; exit 2 ;;
esac
