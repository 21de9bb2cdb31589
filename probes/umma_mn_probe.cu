// umma_mn_probe.cu -- bring-up probe for the 16-bit training backward (round 2):
//   * MN-major A and B operands (SWIZZLE_NONE "interleave" canonical layout) read straight from the
//     "T32" activation layout  [feature/8][32 points][8 features] x 16 bit  that the forward / dgrad
//     epilogues write: D[m = feature of A][n = feature of B] = sum_p A[p][m] B[p][n]   (the wgrad contraction)
//       descriptor: LBO = 128 B (next 8 points, K direction), SBO = 512 B (next 8 features, MN direction)
//   * mixed operand formats in one kind::f16 MMA (A fp16, B bf16): illegal instruction on sm_100a (SNB_PROBE_MIXED=1)
//   * SBO = 0 aliasing (all 16 row groups of A read the same 8 features)
//   * issue-to-completion cycles of MN-major SS MMAs (M128 N256 K16)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probes/umma_mn_probe probes/umma_mn_probe.cu
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

constexpr int kPts = 32;      // points per tile = 2 K16 steps

__host__ __device__ constexpr uint32_t idesc_mn(uint32_t afmt, uint32_t bfmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

struct Args {
  const uint8_t* A;   // T32 tile, FA features: [FA/8][32][8] 16-bit
  const uint8_t* B;   // T32 tile, FB features
  float* D;           // [128][FB]
  long long* cycles;
  int FA, FB, afmt, bfmt, alias, reps;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16 * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < a.FA * kPts * 2 / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(a.A)[i];
  for (int i = tid; i < a.FB * kPts * 2 / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(a.B)[i];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t idesc = idesc_mn((uint32_t)a.afmt, (uint32_t)a.bfmt, 128, (uint32_t)a.FB);
  auto chain = [&]() {
    for (int ks = 0; ks < kPts / 16; ++ks) {
      // K step = 16 points = 256 B along the point axis; LBO (K dir) 128 B, SBO (MN dir) 512 B (0 = alias)
      const uint64_t ad = make_smem_desc(smem_u32(sA) + ks * 256, 128, a.alias ? 0 : 512);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + ks * 256, 128, 512);
      mma_ss(tbase, ad, bd, idesc, ks > 0);
    }
  };
  if (tid == 0) { chain(); mma_commit(&bar); }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < a.FB; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_addr(tbase, warp * 32, c0), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) a.D[tid * a.FB + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < a.reps; ++r) chain();
    mma_commit(&bar);
    mbar_wait(&bar, 1);
    a.cycles[0] = clock64() - t0;
    a.cycles[1] = (long long)a.reps * (kPts / 16);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

static uint16_t f2h(float f) { __half h = __float2half(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }

int run(int FA, int FB, int afmt, int bfmt, int alias, const char* what) {
  std::vector<float> fA(kPts * FA), fB(kPts * FB);
  srand(77 + FA + FB + afmt * 3 + bfmt);
  for (auto& v : fA) v = (float)((rand() % 9) - 4) * 0.5f;
  for (auto& v : fB) v = (float)((rand() % 9) - 4) * 0.25f;
  std::vector<uint8_t> hA((size_t)FA * kPts * 2), hB((size_t)FB * kPts * 2);
  auto put = [&](std::vector<uint8_t>& dst, const std::vector<float>& src, int F, int fmt) {
    for (int p = 0; p < kPts; ++p)
      for (int f = 0; f < F; ++f) {
        const uint16_t u = fmt ? f2bf(src[p * F + f]) : f2h(src[p * F + f]);
        memcpy(&dst[(size_t)(f / 8) * (kPts * 16) + p * 16 + (f % 8) * 2], &u, 2);
      }
  };
  put(hA, fA, FA, afmt);
  put(hB, fB, FB, bfmt);
  uint8_t *dA, *dB;
  float* dD;
  long long* dc;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, 128 * FB * 4)); CK(cudaMalloc(&dc, 16));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * FB * 4));
  Args a{dA, dB, dD, dc, FA, FB, afmt, bfmt, alias, 512};
  const int smem = 48 * 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(a);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * FB);
  long long cyc[2];
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < FB; ++n) {
      const int ma = alias ? (m % 8) : m;
      if (!alias && m >= FA) continue;
      double ref = 0;
      for (int p = 0; p < kPts; ++p) ref += (double)fA[p * FA + ma] * fB[p * FB + n];
      const double e = fabs(ref - D[m * FB + n]);
      if (e > 1e-3) ++bad;
      if (e > maxerr) maxerr = e;
    }
  printf("%-46s FA=%3d FB=%3d: mismatches %6d maxerr %g | %.1f cyc/MMA\n", what, FA, FB, bad, maxerr,
         (double)cyc[0] / (double)cyc[1]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
  return bad ? 1 : 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d\n", prop.name, prop.major, prop.minor);
  int rc = 0;
  rc |= run(128, 256, 0, 0, 0, "MN-major A,B fp16 x fp16 (T32 layout)");
  rc |= run(128, 256, 1, 1, 0, "MN-major A,B bf16 x bf16");
  rc |= run(128, 128, 0, 0, 0, "MN-major fp16, N=128");
  rc |= run(128, 64, 0, 0, 0, "MN-major fp16, N=64");
  rc |= run(128, 32, 0, 0, 0, "MN-major fp16, N=32");
  rc |= run(8, 256, 0, 0, 1, "A aliased with SBO=0 (8 features x16)");
  printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
  // mixed operand formats (A fp16 with B bf16) are NOT accepted by kind::f16 on sm_100a: the launch dies with
  // "an illegal instruction was encountered" (measured, round 2) -- run last, not part of the verdict
  if (getenv("SNB_PROBE_MIXED")) run(128, 256, 0, 1, 0, "MN-major mixed: A fp16, B bf16");
  return rc;
}
