#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/mono_check.py dump gpurun_out/mono_on.npz > gpurun_out/r2_mono_check.txt 2>&1
IPMCMC_LIB=$PWD/gpurun_variants/libipmcmc_nomono.so timeout 600 python tools/mono_check.py dump gpurun_out/mono_off.npz >> gpurun_out/r2_mono_check.txt 2>&1
python tools/mono_check.py compare gpurun_out/mono_on.npz gpurun_out/mono_off.npz >> gpurun_out/r2_mono_check.txt 2>&1
rm -f gpurun_out/mono_on.npz gpurun_out/mono_off.npz
