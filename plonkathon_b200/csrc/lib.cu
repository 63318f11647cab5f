// Single translation unit for libplonk_b200.so (keeps every kernel and its host launcher in one
// module: no relocatable device code needed, and nvcc can inline the field arithmetic everywhere).
#include "ntt.cu"
#include "poly_ops.cu"
#include "msm.cu"
#include "capi.cu"
