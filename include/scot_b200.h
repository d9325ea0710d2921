/*
 * scot_b200.h — C ABI of the B200-native scOT forward/backward engine (libscot_b200.so).
 *
 * This is the drop-in boundary of the hot path named in BASELINE.json: everything behind
 * `ScOT.forward` / autograd backward of the reference (scOT/model.py:1318-1509 and the HuggingFace
 * swinv2 pieces it imports, scOT/model.py:39-47) runs inside the functions declared here.
 * The reference has no FFI of its own (it is pure Python on torch ops); the Python host module
 * poseidon_b200/scOT/model.py binds these entry points with ctypes exactly the way a maintainer of the
 * reference would.
 *
 * Conventions
 *  - every pointer is a raw CUDA device pointer unless the name ends in `_host`;
 *  - the library never allocates, frees or retains device memory: the caller (PyTorch) owns every
 *    buffer, including the workspace/arena handed to the engine;
 *  - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*), no internal
 *    synchronisation, so the calls are CUDA-graph capturable;
 *  - return value 0 = ok, 1 = bad argument / unsupported configuration, 2 = CUDA error;
 *    `scot_last_error()` returns a human readable description (thread local);
 *  - no CPU fallback exists: a missing GPU is an error.
 */
#ifndef SCOT_B200_H_
#define SCOT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCOT_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------------------ */
int scot_abi_version(void);
const char* scot_last_error(void);
/* number of kernels launched by this library since process start (bench.py's `gpu_launches`) */
unsigned long long scot_launch_count(void);

/* ---- GEMM (tcgen05 / TMA) ------------------------------------------------------------------------
 * D[m,n] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulation.
 *   a_mn_major == 0: A(m,k) = A[m*lda + k]   (activations; reduction dimension contiguous)
 *   a_mn_major == 1: A(m,k) = A[k*lda + m]   (wgrad: token dimension is the reduction)
 *   same for B with n.
 * Replaces nn.Linear forward/backward (cuBLAS) at: HF swinv2 query/key/value/dense (:416-418,:531,
 * :571,:586), scOT/model.py:669 (merging), :725-726 (unmerging), :186-190 (ConvNeXt pw convs). */
enum {
  SCOT_EPI_BF16 = 0,        /* out0(bf16) = acc + bias                                               */
  SCOT_EPI_F32 = 1,         /* out0(f32)  = acc + bias                                               */
  SCOT_EPI_GELU = 2,        /* out0(bf16) = h = acc + bias (may be NULL), out1(bf16) = gelu_erf(h)   */
  SCOT_EPI_GELU_BWD = 3,    /* out0(bf16) = acc * gelu_erf'(aux(bf16)); colsum(f32)[n] += column sums */
  SCOT_EPI_RMW_F32 = 4,     /* out0(f32) += acc                                                      */
  SCOT_EPI_ATOMIC_F32 = 5,  /* red.add out0(f32) += acc (split reduction; wgrad)                     */
  SCOT_EPI_ADD_F32_BF16 = 6 /* out0(f32) = acc + bias + aux(f32); out1(bf16, may be NULL) = same     */
};
enum { SCOT_GEMM_TCGEN05 = 0, SCOT_GEMM_SIMT = 1 /* CUDA-core cross-check path used by the tests */ };

typedef struct ScotEpilogue {
  int mode;
  const float* bias; /* [N] or NULL */
  void* out0;
  long ld0;
  void* out1;
  long ld1;
  const void* aux;
  long ldaux;
  float* colsum; /* [N] or NULL */
} ScotEpilogue;

int scot_gemm_bf16(const void* A, long lda, int a_mn_major, const void* B, long ldb, int b_mn_major, int M, int N,
                   int K, const ScotEpilogue* epi, int impl, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCOT_B200_H_ */
