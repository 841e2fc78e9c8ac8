// hostsim_lib.cpp -- builds the CUDA library's own source for the CPU under the emulator in
// cuda_shim.h (TEST INFRASTRUCTURE ONLY; see that header).  Result: tests/hostsim/libsmcdet_hostsim.so
// with the same C ABI as libsmcdet_b200.so but taking HOST pointers.
#define SMC_HOSTSIM 1
#define SMC_HOSTSIM_IMPL 1
#include "../../smcdet_b200/csrc/smcdet_kernels.cu"
