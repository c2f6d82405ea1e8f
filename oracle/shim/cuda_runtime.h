// oracle/shim/cuda_runtime.h -- empty stand-in: src/post_process.hpp:8 includes
// <cuda_runtime.h> but the CPU path (use_gpu=false, paf.cpp:343-344) uses nothing from it.
#pragma once
