# Build the library, the compat test and the host driver for B200 and run them (reference: run.sh:1 built tests/main.cu for sm_86).
set -e
cd "$(dirname "$0")"
ARCH="-gencode arch=compute_100a,code=sm_100a -lineinfo"
nvcc -std=c++17 -O3 $ARCH -shared -Xcompiler -fPIC -o libfa_b200.so kernels/FlashAttention.cu
nvcc -std=c++17 -O2 $ARCH tests/main.cu -o tests/compat_main
nvcc -std=c++17 -O2 $ARCH -x cu main.cpp -o fa_main -L. -lfa_b200 -Xlinker -rpath -Xlinker '$ORIGIN'
./tests/compat_main
./fa_main --props "$@"
