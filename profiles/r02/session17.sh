#!/bin/bash
# round 2, GPU session 17 (1 GPU): timing probes of the 3D kernels (profiles/ab/libprobe.so, built from a patched copy of
# nlps_cellwarp.cu -- profiles/ab/r02_probe*.patch; the product library is untouched)
cd "$(dirname "$0")/../.."
O=gpurun_out/r02_s17; mkdir -p $O
: > $O/probe_phases_v6.txt
for pr in 0 4 8 16 32 64 256 512 128 28 992; do
  NLPS_LIB=$PWD/profiles/ab/libprobe.so timeout 120 python profiles/ab/probe_phases.py 64 $pr >> $O/probe_phases_v6.txt 2>> $O/probe_phases.err
done
cat $O/probe_phases_v6.txt
