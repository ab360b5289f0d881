# Build tuning variants of the library (launch shapes); run them with tools/bench_variants.sh under gpurun.
set -e
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
S=mpc4rl_b200/csrc/rlmpc_b200.cu
build() { name=$1; shift; nvcc $F "$@" -o mpc4rl_b200/variants_$name.so $S & }
build ss64 -DRLMPC_STAGE_TPB=64
build ss3 -DRLMPC_SS_MINB=3
build ss4 -DRLMPC_SS_MINB=4
build qp1m6 -DRLMPC_QP1_MINB=6
build qp1m8 -DRLMPC_QP1_MINB=8
build sw5 -DRLMPC_SW_MINB=5
build sw6 -DRLMPC_SW_MINB=6
build tpb128 -DRLMPC_TPB=128
build lin3 -DRLMPC_LIN_MINB=3
wait
ls -la mpc4rl_b200/variants_*.so
