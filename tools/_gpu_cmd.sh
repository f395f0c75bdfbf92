timeout 100 python tools/prof_small.py --which mfcc,fastconv > gpurun_out/prof_small_r1n.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"mfcc_kernel" -s 1 -c 1 -o gpurun_out/k_mfcc python tools/prof_small.py --which mfcc --iters 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"fastconv_kernel" -s 1 -c 1 -o gpurun_out/k_fastconv python tools/prof_small.py --which fastconv --iters 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"pitch_kernel" -s 1 -c 1 -o gpurun_out/k_pitch python tools/prof_pitch.py --streams 1184 --seconds 8 --iters 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"fft_c2c_big" -s 1 -c 1 -o gpurun_out/k_fft8k python tools/prof_fft.py 8192 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"fft_c2c_big" -s 1 -c 1 -o gpurun_out/k_fft4k python tools/prof_fft.py 4096 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"roundtrip_kernel" -s 1 -c 1 -o gpurun_out/k_rt python tools/bench_extras.py --only roundtrip --quick --out gpurun_out/x2.json > /dev/null 2>&1
ls -la gpurun_out/k_*.ncu-rep | wc -l; cat gpurun_out/prof_small_r1n.log
