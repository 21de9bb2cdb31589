// wgrad_tc.cu -- weight gradients of one nn.Linear on 5th-gen tensor cores.
//
//   dW[n][col_off + k] += sum_p dY[p][n] * X[p][k]      db[n] += sum_p dY[p][n]
//
// (reference: autograd of `nn.Linear` inside models/nerf.py:105-148; the SIMT version of the same
// contraction is wgrad_kernel in field_bwd.cu.)  dY (P, N) and X (P, ldx) are the plain row-major
// fp32 tensors the training forward / the dgrad chain leave in HBM.
//
// As an MMA the reduction runs over POINTS:  D[m = out feature][n' = in feature] += A[m][p] B[n'][p],
// so both operands are needed "points-major" -- the transpose of how they sit in HBM.  Eight
// converter warps do that on the way in: a thread owns one feature column and 8 consecutive points,
// loads them with 8 warp-coalesced 4-byte loads, splits every value into bf16 hi + lo and writes one
// 16-byte row of the canonical K-major (SWIZZLE_NONE) core-matrix layout for each -- conflict-free,
// no second pass.  bf16 (not fp16) because gradients span fp32's exponent range; the 3-product
// split  x*w ~ xh*wh + xl*wh + xh*wl  keeps ~16 mantissa bits per operand (the gradient parity bar is
// 1e-3 per tensor).  One elected lane of a ninth warp issues tcgen05.mma (SS, M = 128, N = Kpad,
// K = 16 points per instruction) into a TMEM accumulator that lives for the CTA's whole point slice;
// at the end the 128 x Kpad block goes TMEM -> smem -> coalesced fp32 atomics (split-P reduction).
//
// grid = (N / 128 out-feature blocks, point slices).  Per point and layer a CTA reads 4*(128 + Kpad)
// bytes; the two out-feature blocks of a slice run side by side, so X is an L2 hit for one of them.
// Roofline: HBM (2 KB per point per 256x256 layer, read once) ~ tensor pipe (3 x 160 cycles per 16
// points per CTA, smem-operand rate); see DESIGN.md.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace snb {
using namespace umma;

namespace {

constexpr int kWgPoints = 64;            // points per pipeline stage (4 MMA K-steps)
constexpr int kWgConvWarps = 8;
constexpr int kWgThreads = (kWgConvWarps + 1) * 32;

template <int KP>
struct WgGeo {
  static constexpr int kABytes = (kWgPoints / 8) * 128 * 16;   // one of {hi, lo} of dY^T: [k8][128][8]
  static constexpr int kBBytes = (kWgPoints / 8) * KP * 16;    // one of {hi, lo} of X^T:  [k8][KP][8]
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = KP == 256 ? 2 : 4;
  static constexpr int kOutBytes = 128 * (KP + 4) * 4;         // epilogue staging, aliases the ring
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kSmemBytes = (kRingBytes > kOutBytes ? kRingBytes : kOutBytes) + 1024;
  static constexpr int kTmemCols = KP < 32 ? 32 : KP;
};

struct WgradTcArgs {
  const float* dY; int ldy;          // (P, ldy); this CTA's out features are columns [128 * blockIdx.x, +128)
  const float* X; int ldx;           // (P, ldx)
  int K;                             // valid columns of X (<= KP)
  float* dW; int ldw; int col_off;   // dW (N, ldw): the block lands at columns [col_off, col_off + K)
  float* db;                         // nullable
  uint32_t* x_pos_bits;              // nullable, KP = 256 only: (P, 8) words, bit c of word w = [X[p][32 w + c] > 0]
  long long P;
  long long rows_per_split;          // multiple of kWgPoints
};

// 8 consecutive points of one feature column -> one 16-byte row of the hi plane and of the lo plane
__device__ __forceinline__ void split8_store(const float (&v)[8], unsigned char* hi_dst, unsigned char* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    const float b0 = __uint_as_float(h[j] << 16), b1 = __uint_as_float(h[j] & 0xffff0000u);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * j] - b0, v[2 * j + 1] - b1);
    l[j] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int KP>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(WgradTcArgs a) {
  using G = WgGeo<KP>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[G::kStages], empty[G::kStages], d_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_off = blockIdx.x * 128;
  const long long r_begin = (long long)blockIdx.y * a.rows_per_split;
  const long long r_end = r_begin + a.rows_per_split < a.P ? r_begin + a.rows_per_split : a.P;
  const int n_stages_total = r_end > r_begin ? (int)((r_end - r_begin + kWgPoints - 1) / kWgPoints) : 0;

  if (tid == 0) {
    for (int i = 0; i < G::kStages; ++i) { mbar_init(&full[i], kWgConvWarps * 32); mbar_init(&empty[i], 1); }
    mbar_init(&d_full, 1);
    fence_mbar_init();
  }
  if (warp == kWgConvWarps) tmem_alloc<G::kTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == kWgConvWarps) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(kFmtBF16, 128, KP);
      for (int it = 0; it < n_stages_total; ++it) {
        const int st = it % G::kStages;
        mbar_wait(&full[st], (it / G::kStages) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(ring + (size_t)st * G::kStageBytes);
        const uint64_t a_hi = make_smem_desc(base, 128 * 16, 128);
        const uint64_t a_lo = make_smem_desc(base + G::kABytes, 128 * 16, 128);
        const uint64_t b_hi = make_smem_desc(base + 2 * G::kABytes, KP * 16, 128);
        const uint64_t b_lo = make_smem_desc(base + 2 * G::kABytes + G::kBBytes, KP * 16, 128);
        constexpr uint64_t kStepA = (2 * 128 * 16) >> 4, kStepB = (2 * KP * 16) >> 4;   // one K16 step, 16-B units
#pragma unroll
        for (int ks = 0; ks < kWgPoints / 16; ++ks) {
          mma_ss(tbase, a_hi + ks * kStepA, b_hi + ks * kStepB, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          mma_ss(tbase, a_lo + ks * kStepA, b_hi + ks * kStepB, idesc, 1u);
          mma_ss(tbase, a_hi + ks * kStepA, b_lo + ks * kStepB, idesc, 1u);
        }
        mma_commit(&empty[st]);
      }
      mma_commit(&d_full);
    }
    __syncwarp();
  } else {
    // ======================= converters =======================
    // A batch = 32 points (half a stage).  Per batch a thread owns 2 groups of dY^T (feature
    // fa = tid & 127, k8 block (tid >> 7) + 2 i) and KP / 64 groups of X^T (group index
    // tid + 256 g -> feature idx % KP, k8 block idx / KP); a group = 8 consecutive points of one
    // feature.  Two register sets: the loads of batch b + 1 are in flight while batch b is split
    // and stored, so HBM latency is paid once per batch, not once per group.
    constexpr int kBatchPts = 32;
    constexpr int kGa = 2, kGb = (KP * (kBatchPts / 8) + 255) / 256;
    const int fa = tid & 127;
    float bias_acc = 0.f;
    const int n_batches = 2 * n_stages_total;
    const float* dy_col = a.dY + n_off + fa;                 // this thread's dY column
    const size_t ldy = (size_t)a.ldy, ldx = (size_t)a.ldx;
    auto load_batch = [&](int bi, float (&va)[kGa][8], float (&vb)[kGb][8]) {
      const long long p0 = r_begin + (long long)bi * kBatchPts;
      const bool full = p0 + kBatchPts <= r_end;               // whole batch in range: no per-load bound checks
#pragma unroll
      for (int i = 0; i < kGa; ++i) {
        const long long p = p0 + ((tid >> 7) + 2 * i) * 8;
        const float* src = dy_col + (size_t)p * ldy;
        if (full) {
#pragma unroll
          for (int r = 0; r < 8; ++r) va[i][r] = __ldg(src + r * ldy);
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) va[i][r] = (p + r < r_end) ? __ldg(src + r * ldy) : 0.f;
        }
      }
#pragma unroll
      for (int g = 0; g < kGb; ++g) {
        const int idx = tid + 256 * g, fb = idx % KP, j = idx / KP;
        const long long p = p0 + j * 8;
        const float* src = a.X + (size_t)p * ldx + fb;
        const bool col_ok = j < kBatchPts / 8 && fb < a.K;
        if (full && col_ok) {
#pragma unroll
          for (int r = 0; r < 8; ++r) vb[g][r] = __ldg(src + r * ldx);
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) vb[g][r] = (col_ok && p + r < r_end) ? __ldg(src + r * ldx) : 0.f;
        }
      }
    };
    auto store_batch = [&](int bi, const float (&va)[kGa][8], const float (&vb)[kGb][8]) {
      const int it = bi >> 1, half = bi & 1, st = it % G::kStages;
      if (half == 0) mbar_wait(&empty[st], ((it / G::kStages) & 1) ^ 1);
      unsigned char* sa_hi = ring + (size_t)st * G::kStageBytes;
      unsigned char* sa_lo = sa_hi + G::kABytes;
      unsigned char* sb_hi = sa_hi + 2 * G::kABytes;
      unsigned char* sb_lo = sb_hi + G::kBBytes;
      const int j0 = half * (kBatchPts / 8);
#pragma unroll
      for (int i = 0; i < kGa; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) bias_acc += va[i][r];
        const int off = (j0 + (tid >> 7) + 2 * i) * (128 * 16) + fa * 16;
        split8_store(va[i], sa_hi + off, sa_lo + off);
      }
#pragma unroll
      for (int g = 0; g < kGb; ++g) {
        const int idx = tid + 256 * g, fb = idx % KP, j = idx / KP;
        if (j < kBatchPts / 8) {
          const int off = (j0 + j) * (KP * 16) + fb * 16;
          split8_store(vb[g], sb_hi + off, sb_lo + off);
        }
        if (KP == 256 && a.x_pos_bits != nullptr && blockIdx.x == 0) {
          // the ReLU mask the following dgrad needs, as a by-product: lanes hold 32 consecutive features of
          // the same 8 points, so one ballot per point is that point's mask word
          const long long p = r_begin + (long long)bi * kBatchPts + j * 8;
          uint32_t mine = 0;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const uint32_t w = __ballot_sync(0xffffffffu, vb[g][r] > 0.f);
            if ((tid & 31) == r) mine = w;
          }
          if ((tid & 31) < 8 && p + (tid & 31) < r_end) a.x_pos_bits[(p + (tid & 31)) * 8 + (fb >> 5)] = mine;
        }
      }
      if (half == 1) {
        fence_proxy_async_smem();     // generic-proxy smem writes -> visible to tcgen05.mma
        mbar_arrive(&full[st]);
      }
    };
    {
      float xa[kGa][8], xb[kGb][8], ya[kGa][8], yb[kGb][8];
      if (n_batches > 0) load_batch(0, xa, xb);
      for (int bi = 0; bi < n_batches; bi += 2) {
        load_batch(bi + 1, ya, yb);               // n_batches is even
        store_batch(bi, xa, xb);
        if (bi + 2 < n_batches) load_batch(bi + 2, xa, xb);
        store_batch(bi + 1, ya, yb);
      }
    }
    if (a.db != nullptr && n_stages_total > 0) atomicAdd(a.db + n_off + fa, bias_acc);

    // ======================= epilogue: TMEM -> smem -> atomics =======================
    if (n_stages_total > 0) {
      mbar_wait(&d_full, 0);
      tc_fence_after();
      float* out = reinterpret_cast<float*>(ring);       // [128][KP + 4]; every MMA has retired, the ring is free
      constexpr int kLd = KP + 4;
      if (warp < 4) {
        const int row = warp * 32 + (tid & 31);
#pragma unroll 1
        for (int c0 = 0; c0 < KP; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(out + row * kLd + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kWgConvWarps * 32) : "memory");
      for (int e = tid; e < 128 * KP; e += kWgConvWarps * 32) {
        const int m = e / KP, k = e - m * KP;
        if (k < a.K) atomicAdd(a.dW + (size_t)(n_off + m) * a.ldw + a.col_off + k, out[m * kLd + k]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWgConvWarps) tmem_dealloc<G::kTmemCols>(tbase);
}

int wg_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

template <int KP>
int launch_wgrad_tc(WgradTcArgs a, int N, cudaStream_t st) {
  using G = WgGeo<KP>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes);
    if (e != cudaSuccess) return fail(SNB_ERR_CUDA, "cudaFuncSetAttribute(wgrad_tc): %s", cudaGetErrorString(e));
    configured = true;
  }
  const int nb = N / 128;
  int splits = wg_sms() / nb;
  if (splits < 1) splits = 1;
  long long rows = (a.P + splits - 1) / splits;
  rows = (rows + kWgPoints - 1) / kWgPoints * kWgPoints;
  splits = (int)((a.P + rows - 1) / rows);
  a.rows_per_split = rows;
  wgrad_tc_kernel<KP><<<dim3(nb, splits), kWgThreads, G::kSmemBytes, st>>>(a);
  return check_launch("wgrad_tc_kernel");
}

}  // namespace

// run_wgrad (field_bwd.cu) on tensor cores: same accumulate-into semantics; optionally also emits the
// sign bits of X (the ReLU mask of the layer's input) for the dgrad that follows
int run_wgrad_tc(const float* dY, int N, const float* X, int ldx, int K, float* dW, int ldw, int col_off, float* db,
                 uint32_t* x_pos_bits, long long P, cudaStream_t st) {
  if (P == 0) return SNB_OK;
  if (N % 128 != 0 || K > 256) return fail(SNB_ERR_INVALID, "run_wgrad_tc: unsupported shape N=%d K=%d", N, K);
  if (x_pos_bits != nullptr && K != 256) return fail(SNB_ERR_INVALID, "run_wgrad_tc: mask bits need K = 256");
  WgradTcArgs a{dY, N, X, ldx, K, dW, ldw, col_off, db, x_pos_bits, P, 0};
  if (K > 64) return launch_wgrad_tc<256>(a, N, st);
  if (K > 32) return launch_wgrad_tc<64>(a, N, st);
  return launch_wgrad_tc<32>(a, N, st);
}

}  // namespace snb
