for L in default 12 24; do
  if [ $L = default ]; then unset JDSP_FFT_FUSED_LOOK; else export JDSP_FFT_FUSED_LOOK=$L; fi
  echo "fused look=$L"; JDSP_FFT_FUSED=1 timeout 200 python tools/sweep_quick.py | tail -3
done
unset JDSP_FFT_FUSED_LOOK
JDSP_FFT_FUSED=1 timeout 200 python -m pytest tests/test_parity.py -x -q -m gpu -k fft 2>&1 | tail -1
for S in 148 592 1184 2368; do
  echo "streams $S tile:"; JDSP_DENOISE_KERNEL=tile timeout 100 python tools/prof_denoise.py --streams $S --seconds 8 | tail -1
  echo "streams $S stream:"; timeout 100 python tools/prof_denoise.py --streams $S --seconds 8 | tail -1
done
