#!/bin/bash
# round 2, GPU session 17 (1 GPU): timing probe -- the 3D kernels without the exponential / without the X loads
# (profiles/ab/libprobe.so is built from a patched copy of nlps_cellwarp.cu: `fexp_poly` and the T.X loads of cw_kin / cw_g2p
# switchable at run time; the product library is untouched)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s17; mkdir -p $O
NLPS_LIB=$PWD/profiles/ab/libprobe.so timeout 600 python profiles/ab/probe_exp.py 64 > $O/probe.json 2> $O/probe.err; echo "rc=$?"; cat $O/probe.json; tail -3 $O/probe.err
