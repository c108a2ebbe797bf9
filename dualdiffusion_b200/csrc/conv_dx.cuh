// Internal interface between the MPConv dispatcher (conv_igemm.cu) and the tap-stacked 3x3 kernel (conv3x3_dx.cu).
#pragma once

#include <cuda_runtime.h>

struct DxConvArgs {
    const void* x;          // bf16 [B][H][W][Cin]
    const void* w;          // bf16 [Cout][9][Cin/groups]  (dd_weight_prep)
    void* out;              // bf16 [B][H][W][Cout]
    int B, H, W, Cin, Cout, groups;
    int epi, epi2;          // DD_EPI_*, DD_EPI2_*
    float alpha, beta, clip;
    const float* scale;     // fp32 [B][Cout]
    const float* scale2;    // fp32 [B][Cout]
    const void* residual;   // bf16 [B][H][W][Cout]
    void* out2;             // bf16 [B][H][W][Cout]
};

// 0 = launched, > 0 = error (dd_last_error set), -1 = shape not handled by this kernel (caller falls back).
int dd_launch_conv3x3_dx(const DxConvArgs& a, cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency); null if unavailable.
void* dd_tensormap_encode_fn();

// Role-timeline buffer of the diagnostic kernel instantiations (DD_CONV_TRACE=1): [4 CTAs][8 slots][64 tiles] clock64
// stamps, read back through dd_conv_trace_read (tools/trace_halo.py); meta = launch geometry of the last traced launch.
unsigned long long* dd_conv_trace_buffer(bool allocate);
int* dd_conv_trace_meta();
