// umma2_probe.cu -- bring-up probe for the 2-CTA (cta_group::2) tcgen05 path + lean MMA-rate timing.
//   * cluster of 2 CTAs; M = 256 (128 rows per CTA), B split by rows across the pair
//   * TMEM alloc/dealloc cta_group::2, tcgen05.mma.cta_group::2 (SS and TS), commit multicast
//   * cross-CTA hand-off: remote mbarrier arrive (mapa + arrive.release.cluster)
//   * timing: issue-lean loops (descriptors precomputed, one elected lane, converged warp) for
//     cta_group::1 N=128 / N=256 and cta_group::2 N=128 / N=256
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probes/umma2_probe probes/umma2_probe.cu
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

// ---- cta_group::2 primitives (candidates for umma.cuh once validated)
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_result)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma2_commit(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}

struct Args {
  const uint16_t* A;   // [256][K]
  const uint16_t* B;   // [N][K]
  float* D;            // [256][N]
  long long* cycles;   // [8]
  int N, K, ts, reps, cg;
};

constexpr uint32_t A_COL = 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_ready;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;               // 128 x K (<=256) 16-bit, canonical [k8][128][16B]   64 KB
  uint8_t* sB = smem + 64 * 1024;   // (N/cgN) x K canonical                                <= 128 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int N = a.N, K = a.K;
  const int cg = a.cg;                 // 1: each CTA runs its own M=128 MMA; 2: one M=256 MMA across the pair
  const int nB = cg == 2 ? N / 2 : N;  // B rows held by this CTA
  const int b_row0 = cg == 2 ? rank * nB : 0;

  if (warp == 0) { if (cg == 2) tmem_alloc2(&tmem_base_s); else tmem_alloc<512>(&tmem_base_s); }
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_ready, cg == 2 ? 256 : 128);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  const uint16_t* Ag = a.A + (size_t)rank * 128 * K;
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e - r * K;
    *reinterpret_cast<uint16_t*>(sA + (k >> 3) * (128 * 16) + r * 16 + (k & 7) * 2) = Ag[e];
  }
  for (int e = tid; e < nB * K; e += 128) {
    const int n = e / K, k = e - n * K;
    *reinterpret_cast<uint16_t*>(sB + (k >> 3) * (nB * 16) + n * 16 + (k & 7) * 2) = a.B[(size_t)(b_row0 + n) * K + k];
  }
  fence_proxy_async_smem();
  if (a.ts) {
    for (int c0 = 0; c0 < K / 2; c0 += 32) {
      uint32_t v[32];
      for (int j = 0; j < 32; ++j) v[j] = (uint32_t)Ag[tid * K + 2 * (c0 + j)] | ((uint32_t)Ag[tid * K + 2 * (c0 + j) + 1] << 16);
      tmem_st32(tmem_addr(tbase, warp * 32, A_COL + c0), v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  if (cg == 2 && rank == 1) mbar_arrive_remote(&bar_ready, 0); else mbar_arrive(&bar_ready);

  const uint32_t idesc = make_idesc(kFmtF16, cg == 2 ? 256 : 128, (uint32_t)N);
  const uint64_t a0 = make_smem_desc(smem_u32(sA), 128 * 16, 128);
  const uint64_t b0 = make_smem_desc(smem_u32(sB), nB * 16, 128);
  const uint32_t stepA = (2 * 128 * 16) >> 4, stepB = (2 * nB * 16) >> 4;
  const bool issuer = (cg == 2) ? (rank == 0) : true;
  const int ksteps = K / 16;

  if (warp == 0 && issuer) {
    if (cg == 2) mbar_wait_cluster(&bar_ready, 0); else mbar_wait(&bar_ready, 0);
    tc_fence_after();
    if (elect_one()) {
      for (int s = 0; s < ksteps; ++s) {
        if (cg == 2) {
          if (a.ts) mma2_ts(tbase, tbase + A_COL + s * 8, b0 + s * stepB, idesc, s > 0);
          else mma2_ss(tbase, a0 + s * stepA, b0 + s * stepB, idesc, s > 0);
        } else {
          if (a.ts) mma_ts(tbase, tbase + A_COL + s * 8, b0 + s * stepB, idesc, s > 0);
          else mma_ss(tbase, a0 + s * stepA, b0 + s * stepB, idesc, s > 0);
        }
      }
      if (cg == 2) mma2_commit(&bar_mma, 3); else mma_commit(&bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_addr(tbase, warp * 32, c0), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) a.D[(size_t)(rank * 128 + tid) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  // ---- lean timing: reps x ksteps MMAs, fully unrolled by 4, one commit
  if (warp == 0 && issuer) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < a.reps; ++r) {
#pragma unroll 4
        for (int s = 0; s < ksteps; ++s) {
          if (cg == 2) {
            if (a.ts) mma2_ts(tbase, tbase + A_COL + s * 8, b0 + s * stepB, idesc, 1);
            else mma2_ss(tbase, a0 + s * stepA, b0 + s * stepB, idesc, 1);
          } else {
            if (a.ts) mma_ts(tbase, tbase + A_COL + s * 8, b0 + s * stepB, idesc, 1);
            else mma_ss(tbase, a0 + s * stepA, b0 + s * stepB, idesc, 1);
          }
        }
      }
      t1 = clock64();
      if (cg == 2) mma2_commit(&bar_mma, 3); else mma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, 1);
    if (elect_one() && rank == 0) {
      const long long t2 = clock64();
      a.cycles[0] = t2 - t0;                       // issue start -> all complete
      a.cycles[1] = (long long)a.reps * ksteps;    // MMAs
      a.cycles[2] = t1 - t0;                       // issue loop only
    }
  } else {
    mbar_wait(&bar_mma, 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { if (cg == 2) tmem_dealloc2(tbase); else tmem_dealloc<512>(tbase); }
}

static uint16_t f2h(float f) {
  __half h = __float2half(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

int run(int cg, int N, int K, int ts, int reps) {
  const int M = 256;
  std::vector<uint16_t> hA((size_t)M * K), hB((size_t)N * K);
  std::vector<float> fA((size_t)M * K), fB((size_t)N * K);
  srand(99 + N + K + cg);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 9) - 4) * 0.5f; hA[i] = f2h(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 9) - 4) * 0.25f; hB[i] = f2h(fB[i]); }
  uint16_t *dA, *dB;
  float* dD;
  long long* dcyc;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMalloc(&dcyc, 64));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, (size_t)M * N * 4));
  CK(cudaMemset(dcyc, 0, 64));
  Args a{dA, dB, dD, dcyc, N, K, ts, reps, cg};
  const int smem = 192 * 1024 + 1024;
  CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe2_kernel<<<2, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("cg=%d N=%3d K=%3d %s: KERNEL ERROR %s\n", cg, N, K, ts ? "TS" : "SS", cudaGetErrorString(e));
    return 3;
  }
  std::vector<float> hD((size_t)M * N);
  long long cyc[8];
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dcyc, 64, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0.f;
      for (int k = 0; k < K; ++k) ref += fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
      const double err = fabs((double)ref - hD[(size_t)m * N + n]);
      if (!(err <= 1e-3)) ++bad;
      if (err > maxerr || err != err) maxerr = err;
    }
  printf("cta_group::%d %s M=%d N=%3d K=%3d: mismatches %6d / %d maxerr %.3g | %lld MMAs: %.1f cyc/MMA to completion, %.1f cyc/MMA issue\n",
         cg, ts ? "TS" : "SS", cg == 2 ? 256 : 128, N, K, bad, M * N, maxerr, cyc[1], cyc[1] ? (double)cyc[0] / cyc[1] : 0.0,
         cyc[1] ? (double)cyc[2] / cyc[1] : 0.0);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dcyc);
  return bad == 0 ? 0 : 1;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d\n", p.name, p.major, p.minor);
  for (int ts = 0; ts < 2; ++ts) {
    run(1, 128, 256, ts, 64);
    run(1, 256, 256, ts, 64);
    run(1, 64, 256, ts, 64);
  }
  for (int ts = 0; ts < 2; ++ts) {
    if (run(2, 128, 64, ts, 64) == 3) return 0;
    run(2, 128, 256, ts, 64);
    run(2, 256, 256, ts, 64);
    run(2, 64, 256, ts, 64);
  }
  return 0;
}
