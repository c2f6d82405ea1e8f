// oracle/shim/cudnn.h -- link-free stand-in for cuDNN so that the reference's
// src/post_process.hpp (peak_finder_t owns a Pool_NCHW_PaddingSame_Max whose ctor
// calls cudnnCreate, src/cudnn_kernel_pool.hpp:13-18) can be constructed on a host
// without a GPU.  The pooling op itself is dead code on the reference's hot path
// (paf.cpp:343-344 passes use_gpu=false); calling it here aborts.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdlib>
typedef enum { CUDNN_STATUS_SUCCESS = 0, CUDNN_STATUS_NOT_SUPPORTED = 9 } cudnnStatus_t;
typedef enum { CUDNN_POOLING_MAX = 0 } cudnnPoolingMode_t;
typedef enum { CUDNN_NOT_PROPAGATE_NAN = 0 } cudnnNanPropagation_t;
typedef enum { CUDNN_TENSOR_NCHW = 0, CUDNN_TENSOR_NHWC = 1 } cudnnTensorFormat_t;
typedef enum { CUDNN_DATA_FLOAT = 0, CUDNN_DATA_DOUBLE = 1 } cudnnDataType_t;
typedef enum { CUDNN_CONVOLUTION = 0, CUDNN_CROSS_CORRELATION = 1 } cudnnConvolutionMode_t;
struct cudnnContext { int unused; };
struct cudnnPoolingStruct { int r, s; };
struct cudnnTensorStruct { cudnnDataType_t dt; int n, c, h, w; };
struct cudnnFilterStruct { int unused; };
struct cudnnConvolutionStruct { int unused; };
typedef cudnnContext* cudnnHandle_t;
typedef cudnnPoolingStruct* cudnnPoolingDescriptor_t;
typedef cudnnTensorStruct* cudnnTensorDescriptor_t;
typedef cudnnFilterStruct* cudnnFilterDescriptor_t;
typedef cudnnConvolutionStruct* cudnnConvolutionDescriptor_t;
inline cudnnStatus_t cudnnCreate(cudnnHandle_t* h) { *h = new cudnnContext{0}; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroy(cudnnHandle_t h) { delete h; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnCreatePoolingDescriptor(cudnnPoolingDescriptor_t* d) { *d = new cudnnPoolingStruct{0, 0}; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyPoolingDescriptor(cudnnPoolingDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnSetPoolingNdDescriptor(cudnnPoolingDescriptor_t d, cudnnPoolingMode_t, cudnnNanPropagation_t, int, const int* w, const int*, const int*) { d->r = w[0]; d->s = w[1]; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnCreateTensorDescriptor(cudnnTensorDescriptor_t* d) { *d = new cudnnTensorStruct{CUDNN_DATA_FLOAT, 0, 0, 0, 0}; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyTensorDescriptor(cudnnTensorDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnSetTensor4dDescriptor(cudnnTensorDescriptor_t d, cudnnTensorFormat_t, cudnnDataType_t dt, int n, int c, int h, int w) { *d = cudnnTensorStruct{dt, n, c, h, w}; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnGetTensor4dDescriptor(cudnnTensorDescriptor_t d, cudnnDataType_t* dt, int* n, int* c, int* h, int* w, int* ns, int* cs, int* hs, int* ws) { *dt = d->dt; *n = d->n; *c = d->c; *h = d->h; *w = d->w; *ws = 1; *hs = d->w; *cs = d->w * d->h; *ns = d->c * d->w * d->h; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnGetPooling2dForwardOutputDim(cudnnPoolingDescriptor_t, cudnnTensorDescriptor_t x, int* n, int* c, int* h, int* w) { *n = x->n; *c = x->c; *h = x->h; *w = x->w; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyFilterDescriptor(cudnnFilterDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnDestroyConvolutionDescriptor(cudnnConvolutionDescriptor_t d) { delete d; return CUDNN_STATUS_SUCCESS; }
inline cudnnStatus_t cudnnPoolingForward(cudnnHandle_t, cudnnPoolingDescriptor_t, const void*, cudnnTensorDescriptor_t, const void*, const void*, cudnnTensorDescriptor_t, void*) { std::abort(); return CUDNN_STATUS_NOT_SUPPORTED; }
