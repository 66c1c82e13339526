#!/bin/bash
# Round-2 visit L: full suite (1-D line analysis, TC mix tweaks), smoke, FFT-vs-DFT record, C1 / C3 bench lines, C2 bench.
tag=${1:-r2l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python tools/fft_vs_dft.py > gpurun_out/${tag}_fft_vs_dft.json 2> gpurun_out/${tag}_fft_vs_dft.err
echo "fft_vs_dft exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_fft_vs_dft.json"))
    for s in d["shapes"]:
        print(" %-42s rfftn %8.1f +gather %8.1f irfftn %8.1f | analysis %8.1f synthesis(+conv,act) %8.1f" % (s["shape"], s["cufft_rfftn_us"], s["cufft_rfftn_plus_corner_gather_us"], s["cufft_irfftn_us"], s["truncated_dft_analysis_us"], s["truncated_dft_synthesis_incl_bypass_conv_bias_gelu_us"]))
except Exception as e:
    print(" parse failed", e)
PY
tail -3 gpurun_out/${tag}_fft_vs_dft.err
for w in burgers1d ns2d_w64 darcy2d; do
  timeout 600 python bench.py --workload $w --steps 100 --no-secondary --cpu-steps 1 > gpurun_out/${tag}_bench_$w.log 2>&1
  echo "bench $w exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$w.log").read().strip().splitlines()[-1])
    print(" value %.1f ms %.3f layer_us %.1f frac %.3f vs_torch %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["layer_fwd_bwd"]["us"], d["roofline"]["layer_fwd_bwd"]["frac"], d.get("vs_torch_gpu") or 0))
    print(" kernels", d["roofline"]["kernels_us"])
except Exception as e:
    print(" parse failed", e)
PY
done
