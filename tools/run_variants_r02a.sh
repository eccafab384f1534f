set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for v in R0 R3 R7 R15 R15f R31; do
  PIHM_B200_LIB=build_exp/$v/libpihm_b200.so timeout 300 python tools/rhs_parity_probe.py 1M 2>&1 | tail -3
done
for v in R15w10 R15w12 R15w6m3; do
  PARITY=0 PIHM_B200_LIB=build_exp/$v/libpihm_b200.so timeout 300 python tools/rhs_parity_probe.py 1M 2>&1 | tail -1
done
for v in R0 R15 R31; do
  PIHM_B200_LIB=build_exp/$v/libpihm_b200.so timeout 300 python tools/rhs_parity_probe.py 1M fbr 2>&1 | tail -3
done
for v in R0 R15; do
  PIHM_B200_LIB=build_exp/$v/libpihm_b200.so timeout 300 python tools/rhs_parity_probe.py 100k 2>&1 | tail -3
done
