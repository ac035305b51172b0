// of_api.cu -- the C ABI and its drivers (host-buffer pipeline, pyramidal level loop, row-band driver) -- compiled
// for the CPU on top of cuda_on_host.h and fake_cudart.cpp: together with the other emul_*.cpp files this gives a
// library with the product's exported symbols whose "device" is the CPU.  TEST INFRASTRUCTURE ONLY.
#include "cuda_on_host.h"

#include "of_api.cu"
