#!/usr/bin/env bash
# round 2: MFCC and fast-conv kernel cycles back to back
bash tools/gpu_r2_kernel.sh mf2 "mfcc or host_forms" mfcc mfcc mfcc_
bash tools/gpu_r2_kernel.sh fc2 "fastconv" fastconv fastconv fastconv_
