nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
P() { PIHM_B200_LIB=build_exp/$1/libpihm_b200.so timeout 300 python tools/rhs_parity_probe.py "${@:2}" 2>&1 | grep "^\[" | cut -c1-230; }
P F0 1M; P F3 1M; PIHM_B200_NO_YSTAGE=1 P F3 1M; P F3late 1M; P F11 1M; PARITY=0 P F3w10 1M
P F0 1M fbr; P F3 1M fbr; P F3f8 1M fbr; P F11f8 1M fbr
P F3 100k; P F3 100k fbr
python -m pytest tests/test_rhs_gpu.py -x -q 2>&1 | tail -5
