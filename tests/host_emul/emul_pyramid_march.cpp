// pyramid_march.cu (the marching Gaussian-pyramid kernel and its launcher) compiled for the CPU; linked with
// emul_pyramid.cpp.  TEST INFRASTRUCTURE.
#include "cuda_on_host.h"

#include "pyramid_march.cu"
