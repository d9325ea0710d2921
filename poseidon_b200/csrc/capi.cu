// extern "C" surface of libscot_b200.so (see include/scot_b200.h) + error / launch-count plumbing.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "internal.h"
#include "scot_b200.h"

static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void scot_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void scot_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" {

int scot_abi_version(void) { return SCOT_ABI_VERSION; }
const char* scot_last_error(void) { return g_err; }
unsigned long long scot_launch_count(void) { return g_launches.load(); }

int scot_gemm_bf16(const void* A, long lda, int a_mn_major, const void* B, long ldb, int b_mn_major, int M, int N,
                   int K, const ScotEpilogue* epi, int impl, void* stream) {
  return scot_gemm_launch(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, epi, impl, (cudaStream_t)stream);
}

}  // extern "C"
