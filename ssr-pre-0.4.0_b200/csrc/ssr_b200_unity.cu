// Unity translation unit: the kernels (kernels.cuh and the headers it includes) are defined once and every
// host file sees the same instantiations (one cubin, one set of symbols).
#include "engine.cu"
#include "gather.cu"
#include "hostgather.cu"
#include "renderer.cu"
#include "capi.cu"
#include "nonuniform.cu"
