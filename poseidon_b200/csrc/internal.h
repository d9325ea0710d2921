// Cross-translation-unit declarations of the launch functions behind the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "scot_b200.h"

int scot_gemm_launch(const void* A, long lda, int a_mn_major, const void* B, long ldb, int b_mn_major, int M, int N,
                     int K, const ScotEpilogue* e, int impl, cudaStream_t stream);
