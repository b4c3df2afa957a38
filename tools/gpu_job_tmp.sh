timeout 60 tools/microbench/exploop 2>&1 | tail -30 > gpurun_out/r2_exploop_chain.log; cat gpurun_out/r2_exploop_chain.log
for cfg in "LN3_FMHA_CHAIN=0" "LN3_FMHA_CHAIN=1" "LN3_FMHA_CHAIN=1 LN3_FMHA_ROTA=1"; do env $cfg timeout 150 python tools/gpu_check_fmha.py 2>&1 | tail -3; done
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x -k "gemm" 2>&1 | tail -4
for bn in 256 0; do LN3_GEMM_BN=$bn timeout 100 python tools/gpu_bench_ops.py 2>&1 | tail -1; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_ops.json')); print({k: round(v['us'],1) for k,v in d.items() if 'gemm' in k})
PY
done
