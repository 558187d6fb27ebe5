// Internal (non-exported) interfaces between the translation units of libshb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace shb {

// tcgen05 gather-GEMM (shb_spiralconv_umma.cu): bf16 storage, gather width Cs in {16,32,64,128}.
bool umma_gather_gemm_supported(int Cs, int Cd, int S);
int umma_gather_gemm(const void* src, const int32_t* table, const int32_t* list, const void* w, const float* bias,
                     void* dst, int B, int rows_src, int rows_dst, int S, int Cs, int Cd, int act, int zero_last,
                     int skip_last, bool sum_mode, cudaStream_t st);

// SHB_DISABLE_UMMA=1 forces the CUDA-core kernels in bf16 mode too (A/B testing)
bool umma_enabled();

}  // namespace shb
