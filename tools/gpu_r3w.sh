#!/bin/bash
# Session-3 visit W: staged slabs as several concurrent bulk copies (DENO_AN_PIECE) - A/B.
tag=${1:-r3w}
mkdir -p gpurun_out
timeout 900 python tools/ab_env.py "DENO_AN_PIECE=0" "DENO_AN_PIECE=8192" "DENO_AN_PIECE=4096" "DENO_AN_PIECE=2048" -- C2 C5 > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
