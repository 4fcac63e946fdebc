B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"])'
echo default; $B 2>/dev/null | python -c "$P"
echo wide G4; FFTCONV_CUDA_N2048_FORCE_WIDE=1 $B 2>/dev/null | python -c "$P"
echo wide G3; FFTCONV_CUDA_N2048_FORCE_WIDE=1 FFTCONV_CUDA_N2048_GROUPS=3 $B 2>/dev/null | python -c "$P"
echo narrow G3 norecomp; FFTCONV_CUDA_N2048_GROUPS=3 $B 2>/dev/null | python -c "$P"
echo default; $B 2>/dev/null | python -c "$P"
