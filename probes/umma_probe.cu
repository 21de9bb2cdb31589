// umma_probe.cu -- hardware bring-up probe for the tcgen05 path (run on a B200 via gpurun).
// Validates, against a CPU reference with exactly-representable inputs:
//   * TMEM alloc + tcgen05.st/ld round trip
//   * SS-mode MMA with the SWIZZLE_NONE K-major canonical smem layout (LBO/SBO roles)
//   * TS-mode MMA (A operand from TMEM, 16-bit pairs packed per 32-bit column)
//   * cp.async.bulk (1-D TMA) + mbarrier complete_tx as the weight-staging path
//   * tcgen05.commit -> mbarrier
// and measures issue-to-completion cycles of back-to-back MMAs (tensor-pipe rate per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probes/umma_probe probes/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

enum { V_SWAP_LBO_SBO = 1, V_TS = 2, V_PACK_SWAP = 4, V_BULK_B = 8 };

struct Args {
  const uint16_t* A;        // [128][K] row-major 16-bit
  const uint16_t* B;        // [N][K] row-major 16-bit
  const uint8_t* Bcanon;    // canonical image of B (host packed)
  float* D;                 // [128][N]
  uint32_t* roundtrip_bad;  // st/ld mismatches
  long long* cycles;        // [2]: total cycles, number of MMAs
  int N, K, variant, fmt, reps;
};

constexpr int M = 128;
constexpr uint32_t A_COL = 256;  // TMEM column base of the A operand (D occupies [0,256))

__global__ void __launch_bounds__(128, 1) probe_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_b;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;                  // up to 128*256*2 = 64 KB
  uint8_t* sB = smem + 64 * 1024;      // up to 256*256*2 = 128 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, K = a.K;

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_b, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  // ---- TMEM st/ld round trip on columns [0,32)
  {
    uint32_t v[32], r[32];
    for (int j = 0; j < 32; ++j) v[j] = (uint32_t)(tid * 1000 + j);
    tmem_st32(tmem_addr(tbase, warp * 32, 0), v);
    tmem_wait_st();
    tmem_ld32(tmem_addr(tbase, warp * 32, 0), r);
    tmem_wait_ld();
    uint32_t bad = 0;
    for (int j = 0; j < 32; ++j) bad += (r[j] != v[j]);
    if (bad) atomicAdd(a.roundtrip_bad, bad);
  }

  // ---- stage operands into the canonical layout: byte offset(r,k) = (k/8)*(R*16) + r*16 + (k%8)*2
  for (int e = tid; e < M * K; e += 128) {
    const int r = e / K, k = e - r * K;
    *reinterpret_cast<uint16_t*>(sA + (k >> 3) * (M * 16) + r * 16 + (k & 7) * 2) = a.A[e];
  }
  if (a.variant & V_BULK_B) {
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(N * K * 2);
      mbar_arrive_expect_tx(&bar_b, bytes);
      for (uint32_t off = 0; off < bytes; off += 16384)
        bulk_g2s(sB + off, a.Bcanon + off, bytes - off < 16384 ? bytes - off : 16384, &bar_b);
    }
    mbar_wait(&bar_b, 0);
  } else {
    for (int e = tid; e < N * K; e += 128) {
      const int n = e / K, k = e - n * K;
      *reinterpret_cast<uint16_t*>(sB + (k >> 3) * (N * 16) + n * 16 + (k & 7) * 2) = a.B[e];
    }
  }
  fence_proxy_async_smem();

  // ---- TS mode: this thread's A row -> TMEM, two K elements per 32-bit column
  if (a.variant & V_TS) {
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t v[32];
      for (int j = 0; j < 32; ++j) {
        const uint32_t e0 = a.A[tid * K + 2 * (c0 + j)], e1 = a.A[tid * K + 2 * (c0 + j) + 1];
        v[j] = (a.variant & V_PACK_SWAP) ? (e1 | (e0 << 16)) : (e0 | (e1 << 16));
      }
      tmem_st32(tmem_addr(tbase, warp * 32, A_COL + c0), v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t idesc = make_idesc((uint32_t)a.fmt, M, (uint32_t)N);
  const bool swap = a.variant & V_SWAP_LBO_SBO;
  auto issue_chain = [&]() {
    for (int s = 0; s < K / 16; ++s) {
      const uint32_t a_addr = smem_u32(sA) + (2 * s) * (M * 16);
      const uint32_t b_addr = smem_u32(sB) + (2 * s) * (N * 16);
      const uint64_t bd = swap ? make_smem_desc(b_addr, 128, N * 16) : make_smem_desc(b_addr, N * 16, 128);
      if (a.variant & V_TS) {
        mma_ts(tbase, tbase + A_COL + s * 8, bd, idesc, s > 0);
      } else {
        const uint64_t ad = swap ? make_smem_desc(a_addr, 128, M * 16) : make_smem_desc(a_addr, M * 16, 128);
        mma_ss(tbase, ad, bd, idesc, s > 0);
      }
    }
  };

  if (tid == 0) {
    issue_chain();
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  // ---- read D back: thread = row
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_addr(tbase, warp * 32, c0), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) a.D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- timing: reps chains back to back, one commit at the end
  if (tid == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < a.reps; ++r) issue_chain();
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 1);
    const long long t1 = clock64();
    a.cycles[0] = t1 - t0;
    a.cycles[1] = (long long)a.reps * (K / 16);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

static uint16_t f2h(float f) {  // small exact integers / halves only
  __half h = __float2half(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);
}

int run(int N, int K, int variant, int fmt, int reps) {
  std::vector<uint16_t> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  srand(1234 + N + K);
  for (int i = 0; i < M * K; ++i) { fA[i] = (float)((rand() % 9) - 4) * 0.5f; hA[i] = fmt ? f2bf(fA[i]) : f2h(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = (float)((rand() % 9) - 4) * 0.25f; hB[i] = fmt ? f2bf(fB[i]) : f2h(fB[i]); }
  std::vector<uint8_t> hBc((size_t)N * K * 2);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) memcpy(&hBc[(size_t)(k >> 3) * (N * 16) + n * 16 + (k & 7) * 2], &hB[n * K + k], 2);
  uint16_t *dA, *dB;
  uint8_t* dBc;
  float* dD;
  uint32_t* dbad;
  long long* dcyc;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dBc, hBc.size()));
  CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMalloc(&dbad, 4));
  CK(cudaMalloc(&dcyc, 16));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBc, hBc.data(), hBc.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, (size_t)M * N * 4));
  CK(cudaMemset(dbad, 0, 4));
  CK(cudaMemset(dcyc, 0, 16));
  Args a{dA, dB, dBc, dD, dbad, dcyc, N, K, variant, fmt, reps};
  const int smem = 192 * 1024 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("variant %2d N=%3d K=%3d fmt=%d: KERNEL ERROR %s\n", variant, N, K, fmt, cudaGetErrorString(e));
    return 3;
  }
  std::vector<float> hD((size_t)M * N);
  uint32_t bad_rt;
  long long cyc[2];
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&bad_rt, dbad, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0.f;
      for (int k = 0; k < K; ++k) ref += fA[m * K + k] * fB[n * K + k];
      const double err = fabs((double)ref - hD[(size_t)m * N + n]);
      if (!(err <= 1e-3)) ++bad;
      if (err > maxerr || err != err) maxerr = err;
    }
  printf("variant %2d [%s%s%s%s] N=%3d K=%3d fmt=%s: mismatches %6d / %d  maxerr %.4g  roundtrip_bad %u  "
         "cycles/MMA %.1f (%lld MMAs)\n",
         variant, (variant & V_TS) ? "TS" : "SS", (variant & V_SWAP_LBO_SBO) ? ",swapLBO" : "",
         (variant & V_PACK_SWAP) ? ",packswap" : "", (variant & V_BULK_B) ? ",bulkB" : "", N, K, fmt ? "bf16" : "f16",
         bad, M * N, maxerr, bad_rt, cyc[1] ? (double)cyc[0] / (double)cyc[1] : 0.0, cyc[1]);
  cudaFree(dA); cudaFree(dB); cudaFree(dBc); cudaFree(dD); cudaFree(dbad); cudaFree(dcyc);
  return bad == 0 ? 0 : 1;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d, %d SMs, clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
  int rc = 0;
  const int variants[] = {0, 8, 2, 2 | 8};   // (swapped LBO/SBO faults: the roles in umma.cuh are right)
  for (int v : variants) rc |= run(256, 64, v, 0, 64) << 0;
  printf("--- shapes / formats with the baseline variants\n");
  run(256, 256, 0, 0, 64);
  run(256, 256, V_TS | V_BULK_B, 0, 64);
  run(128, 256, V_TS | V_BULK_B, 0, 64);
  run(128, 32, V_TS | V_BULK_B, 0, 64);
  run(256, 256, V_TS | V_BULK_B, 1, 64);
  run(256, 256, V_BULK_B, 1, 64);
  run(128, 256, V_BULK_B, 0, 64);
  run(64, 256, V_TS | V_BULK_B, 0, 64);
  printf("--- packswap (expected to mismatch)\n");
  run(256, 64, V_TS | V_PACK_SWAP, 0, 8);
  return 0;
}
