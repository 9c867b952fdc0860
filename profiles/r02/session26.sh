#!/bin/bash
# round 2, GPU session 26 (1 GPU): quick check of the default bench line after the traffic record was re-keyed by kernel machine
# code (5 timed steps, no CPU leg, no e2e) -> profiles/bench/r02_c3_n1_quick_s26.json: 42.2 ms/step, roofline.traffic present.
# (Taken on the build of commit f4283a2; the C3 kernels of the final build are bit-identical, profiles/r02/sass_identity_5867eec.txt
# and the stamp check of bench.py:ncu_traffic.)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r02_s26
timeout 180 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_s26/bench_c3_quick.json 2> gpurun_out/r02_s26/bench_c3_quick.err; echo rc=$?
