// Internal (non-exported) interfaces between the translation units of libshb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shb {

// tcgen05 gather-GEMM (shb_spiralconv_umma.cu): bf16 storage, gather width Cs in {16,32,64,128}.
bool umma_gather_gemm_supported(int Cs, int Cd, int S);
int umma_gather_gemm(const void* src, const int32_t* table, const int32_t* keyptr, const int32_t* list, const void* w,
                     const float* bias,
                     void* dst, int B, int rows_src, int rows_dst, int S, int Cs, int Cd, int act, int zero_last,
                     int skip_last, bool sum_mode, cudaStream_t st);

// tcgen05 weight gradient: the whole (K x Cout) accumulator resident in TMEM; fp32 gw/gb out.
bool umma_wgrad_supported(int Cin, int Cout, int S);
size_t umma_wgrad_workspace(int B, int rows_out, int S, int Cin, int Cout);
int umma_wgrad(const void* x, const int32_t* table, const void* gz, float* gw, float* gb, void* workspace, int B,
               int rows_in, int rows_out, int S, int Cin, int Cout, int src_dummy_zero, cudaStream_t st);

// debug: device buffer of 3*4*512 int64 receiving CTA 0's per-stage clock64 stamps (null = off)
void umma_set_trace(long long* buf);

// SHB_DISABLE_UMMA=1 forces the CUDA-core kernels in bf16 mode too (A/B testing)
bool umma_enabled();

}  // namespace shb
