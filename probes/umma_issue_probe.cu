// umma_issue_probe.cu -- how deep is the tcgen05.mma queue and what does issuing cost?
//   cluster of 2 CTAs, cta_group::2, TS operands (A in TMEM), M = 256, N = 128 or 256, K16 steps.
//   burst<n>: from an idle pipe, one elected lane issues n MMAs back to back, then one commit;
//             reports cycles for the issue loop, for the commit instruction and until completion.
//   chunked<n>: R rounds of [n MMAs + 2 commits + a G-cycle issuer gap]: the issuer pattern of
//             field_tc_kernel, to see how long a gap the queue hides.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probes/umma_issue_probe probes/umma_issue_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../sinnerf_b200/csrc/umma.cuh"

using namespace snb::umma;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

constexpr uint32_t A_COL = 256;

struct Args {
  long long* out;   // [rows][4]
  int N;            // 128 or 256
  int gap;          // issuer gap between chunks (cycles), chunked mode
  int rounds;
};

template <int n>
__device__ __forceinline__ void issue_n(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t stepB) {
#pragma unroll
  for (int s = 0; s < n; ++s) mma2_ts_lohi(d, a + (s & 15) * 8, b_lo + (s & 15) * stepB, b_hi, idesc, 1);
}

template <int n>
__device__ __forceinline__ void burst(long long* out, uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                      uint32_t stepB, uint64_t* bar, uint32_t& phase) {
  long long t0 = 0, t1 = 0, t2 = 0;
  if (elect_one()) {
    t0 = clock64();
    issue_n<n>(d, a, b_lo, b_hi, idesc, stepB);
    t1 = clock64();
    mma2_commit(bar);
    t2 = clock64();
  }
  __syncwarp();
  mbar_wait(bar, phase);
  phase ^= 1;
  tc_fence_after();
  if (elect_one()) {
    const long long t3 = clock64();
    out[0] = n; out[1] = t1 - t0; out[2] = t2 - t1; out[3] = t3 - t0;
  }
  __syncwarp();
}

template <int n>
__device__ __forceinline__ void chunked(long long* out, const Args& a, uint32_t d, uint32_t acol, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t stepB, uint64_t* bar, uint64_t* bar2, uint32_t& phase) {
  long long t0 = clock64();
  for (int r = 0; r < a.rounds; ++r) {
    if (elect_one()) {
      issue_n<n>(d, acol, b_lo, b_hi, idesc, stepB);
      mma2_commit(bar2);                 // stands for the ring "empty" commit: nobody waits on it
      if (r == a.rounds - 1) mma2_commit(bar);
    }
    __syncwarp();
    const long long g0 = clock64();
    while (clock64() - g0 < a.gap) {}
  }
  mbar_wait(bar, phase);
  phase ^= 1;
  tc_fence_after();
  if (elect_one()) {
    const long long t3 = clock64();
    out[0] = n; out[1] = a.gap; out[2] = a.rounds; out[3] = t3 - t0;
  }
  __syncwarp();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int nB = a.N / 2;
  if (warp == 0) tmem_alloc_pair(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
  for (int e = tid; e < 32 * 1024 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = 0;
    for (int c0 = 0; c0 < 128; c0 += 32) tmem_st32(tbase + ((uint32_t)(warp * 32) << 16) + A_COL + c0, v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  const uint32_t idesc = make_idesc(kFmtF16, 256, (uint32_t)a.N);
  const uint64_t b0 = make_smem_desc(smem_u32(smem), nB * 16, 128);
  const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
  const uint32_t stepB = (2 * nB * 16) >> 4;
  if (warp == 0 && rank == 0) {
    uint32_t phase = 0;
    long long* o = a.out;
    const uint32_t acol = tbase + A_COL;
    burst<1>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase);   // warm-up
    burst<1>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    burst<2>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    burst<4>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    burst<8>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    burst<16>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    burst<24>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    burst<32>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    burst<64>(o, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, phase); o += 4;
    chunked<8>(o, a, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, &bar2, phase); o += 4;
    chunked<24>(o, a, tbase, acol, b_lo, b_hi, idesc, stepB, &bar, &bar2, phase); o += 4;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tbase);
}

int run(int N, int gap, bool print_burst) {
  long long* d;
  CK(cudaMalloc(&d, 64 * 8));
  CK(cudaMemset(d, 0, 64 * 8));
  Args a{d, N, gap, 64};
  const int smem = 64 * 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<2, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("KERNEL ERROR %s\n", cudaGetErrorString(e)); return 3; }
  long long h[64];
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  if (print_burst)
    for (int i = 0; i < 8; ++i)
      printf("N=%3d burst of %2lld MMAs from idle: issue %5lld cyc (%.1f/MMA)  commit %3lld cyc  done after %5lld cyc\n", N, h[i * 4],
             h[i * 4 + 1], (double)h[i * 4 + 1] / h[i * 4], h[i * 4 + 2], h[i * 4 + 3]);
  for (int i = 8; i < 10; ++i) {
    const double per = (double)h[i * 4 + 3] / h[i * 4 + 2];
    printf("N=%3d chunks of %2lld MMAs + commit, issuer gap %4lld cyc: %.0f cyc/chunk (pure MMA %.0f)\n", N, h[i * 4], h[i * 4 + 1], per,
           (double)h[i * 4] * (N == 128 ? 64.4 : 128.3));
  }
  cudaFree(d);
  return 0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d\n", p.name, p.major, p.minor);
  for (int N : {128, 256}) {
    bool first = true;
    for (int gap : {0, 100, 200, 300, 400, 600}) { if (run(N, gap, first)) return 1; first = false; }
  }
  return 0;
}
