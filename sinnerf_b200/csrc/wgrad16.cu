// wgrad16.cu -- weight gradients of one nn.Linear from 16-bit T32 operands (act16.cuh), on tensor cores.
//
//   dW[n][col_off + k] += (1 / scale) * sum_p dY[p][n] X[p][k]        db[n] += (1 / scale) * sum_p dY[p][n]
//   optional head rows:  dH[r][k] += (1 / scale2) * sum_p (hg[p][r] + hg[p][r + 4]) X[p][k],  r < 4   (sigma / rgb head
//                        weights; features r + 4 of the hg cell hold the fp16 rounding residual of feature r)
//
// (reference: autograd of `nn.Linear` inside models/nerf.py:105-148.)  dY, X and hg are fp16 tensors in
// the T32 layout: a 32-point tile copied verbatim into shared memory IS the MN-major SWIZZLE_NONE canonical
// operand of tcgen05.mma (probes/umma_mn_probe.cu), so the whole kernel is
//   producer (1 elected thread)   cp.async.bulk of the tile's dY block (FA x 64 B) and X block (FB x 64 B)
//                                 into a kStages-deep ring (mbarrier complete_tx);
//   issuer   (1 elected thread)   per tile 2 K-steps (K = 16 points) x NM out-feature blocks of
//                                 tcgen05.mma SS  M = 128, N = FB  into TMEM accumulators that live for the
//                                 CTA's whole slice of points; one product (fp16 x fp16, fp32 accumulate);
//   bias / epilogue (4 warps)     column sums of the dY tile straight from shared memory (lane = point,
//                                 8 features per 16-byte cell) while the MMAs run; at the end TMEM ->
//                                 smem -> scaled, coalesced fp32 vector atomics (split-P reduction).
// No converter warps, no transposition, every HBM byte read once: 2 (FA + FB) bytes per point
// (1 KB for a 256 x 256 layer; the fp32 version moved 2 KB and converted all of it in registers).
// Roofline: HBM.  Shared-memory traffic per tile (TMA write + 2 NM operand reads + bias read) is the second
// limit, the tensor pipe (2 NM MMAs of ~160 cycles per 32 points) the third.
#include "act16.cuh"
#include "common.cuh"
#include "umma.cuh"

namespace snb {
using namespace umma;

namespace {

constexpr int kW16EpiWarps = 4;
constexpr int kW16LoadWarp = kW16EpiWarps, kW16MmaWarp = kW16EpiWarps + 1;
constexpr int kW16Threads = (kW16EpiWarps + 2) * 32;

// NM: 128-row out-feature blocks of dY (FA = 128 NM; 0 = head rows only); FB: X features (MMA N); kHead: hg operand
template <int NM, int FB, bool kHead>
struct W16Geo {
  static constexpr int kFA = 128 * NM;
  static constexpr int kDyBytes = kFA * 64, kXBytes = FB * 64, kHgBytes = kHead ? 512 : 0;
  static constexpr int kStageBytes = kDyBytes + kXBytes + kHgBytes;
  static constexpr int kStagesRaw = (160 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kOutLd = FB + 4;
  static constexpr int kOutBytes = 128 * kOutLd * 4;                   // epilogue staging, aliases the ring
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kSmemBytes = (kRingBytes > kOutBytes ? kRingBytes : kOutBytes) + 1024;
  static constexpr int kAccCols = (NM + (kHead ? 1 : 0)) * FB;
  static constexpr int kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
  static_assert(kAccCols <= 512, "TMEM columns");
  static_assert(kStages >= 2 && kSmemBytes <= 227 * 1024, "shared memory");
  static_assert(FB % 16 == 0 && FB >= 16 && FB <= 256, "MMA N");
};

struct W16Args {
  const unsigned char* dY;     // (Ppad, FA) fp16 T32 (unused when NM == 0)
  const unsigned char* X;      // (Ppad, FB) fp16 T32
  const unsigned char* hg;     // (Ppad, 8)  fp16 T32 (kHead)
  int K;                       // valid columns of X (<= FB)
  float* dW; int ldw; int col_off;
  float* db;                   // nullable
  const float* scale;          // device: dY is stored as true * (*scale)
  float* dH[8]; int ldh;       // kHead: destination row pointers (nullable per row), row stride unused (rows are separate tensors)
  const float* scale2;         // device: scale of hg
  long long n_tiles;           // 32-point tiles (Ppad / 32)
  long long tiles_per_cta;
};

__host__ __device__ constexpr uint32_t idesc_mn_f16(uint32_t M, uint32_t N) {
  // kind::f16, fp16 x fp16 -> fp32, A and B MN-major
  return (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <int NM, int FB, bool kHead>
__global__ void __launch_bounds__(kW16Threads, 1) wgrad16_kernel(W16Args a) {
  using G = W16Geo<NM, FB, kHead>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[G::kStages], empty[G::kStages], d_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_begin = (long long)blockIdx.x * a.tiles_per_cta;
  const long long t_end = t_begin + a.tiles_per_cta < a.n_tiles ? t_begin + a.tiles_per_cta : a.n_tiles;
  const int n_my = t_end > t_begin ? (int)(t_end - t_begin) : 0;

  if (tid == 0) {
    for (int i = 0; i < G::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + (NM > 0 ? kW16EpiWarps * 32 : 0)); }
    mbar_init(&d_full, 1);
    fence_mbar_init();
  }
  if (warp == kW16MmaWarp) tmem_alloc<G::kTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == kW16LoadWarp) {
    // ======================= producer: one tile = two (three) contiguous runs in HBM =======================
    if (elect_one()) {
      for (int i = 0; i < n_my; ++i) {
        const int st = i % G::kStages;
        mbar_wait(&empty[st], ((i / G::kStages) & 1) ^ 1);
        unsigned char* dst = ring + (size_t)st * G::kStageBytes;
        const long long t = t_begin + i;
        mbar_arrive_expect_tx(&full[st], (uint32_t)G::kStageBytes);
        if (NM > 0) {
          const unsigned char* src = a.dY + (size_t)t * G::kDyBytes;
          for (int o = 0; o < G::kDyBytes; o += 16384) bulk_g2s(dst + o, src + o, G::kDyBytes - o < 16384 ? G::kDyBytes - o : 16384, &full[st]);
        }
        {
          const unsigned char* src = a.X + (size_t)t * G::kXBytes;
          for (int o = 0; o < G::kXBytes; o += 16384)
            bulk_g2s(dst + G::kDyBytes + o, src + o, G::kXBytes - o < 16384 ? G::kXBytes - o : 16384, &full[st]);
        }
        if (kHead) bulk_g2s(dst + G::kDyBytes + G::kXBytes, a.hg + (size_t)t * 512, 512, &full[st]);
      }
    }
    __syncwarp();
  } else if (warp == kW16MmaWarp) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      constexpr uint32_t idesc = idesc_mn_f16(128, FB);
      for (int i = 0; i < n_my; ++i) {
        const int st = i % G::kStages;
        mbar_wait(&full[st], (i / G::kStages) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(ring + (size_t)st * G::kStageBytes);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          // K step = 16 points = 256 B along the point axis; LBO (next 8 points) 128 B, SBO (next 8 features) 512 B
          const uint64_t bd = make_smem_desc(base + G::kDyBytes + ks * 256, 128, 512);
          const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
          for (int mb = 0; mb < NM; ++mb) {
            const uint64_t ad = make_smem_desc(base + mb * (128 * 64) + ks * 256, 128, 512);
            mma_ss(tbase + mb * FB, ad, bd, idesc, acc);
          }
          if (kHead) {
            // the 8 head-gradient features as all 16 row groups of A (SBO = 0): rows 8..127 of the result repeat rows 0..7
            const uint64_t ad = make_smem_desc(base + G::kDyBytes + G::kXBytes + ks * 256, 128, 0);
            mma_ss(tbase + NM * FB, ad, bd, idesc, acc);
          }
        }
        mma_commit(&empty[st]);
      }
      mma_commit(&d_full);
    }
    __syncwarp();
  } else {
    // ======================= bias: column sums of dY from the staged tiles =======================
    // warp w owns feature groups [w * kGw, +kGw); lane = point.  acc[g][j] = partial sum over this lane's points.
    constexpr int kGw = NM > 0 ? (G::kFA / 8) / kW16EpiWarps : 1;
    float acc[kGw][8];
#pragma unroll
    for (int g = 0; g < kGw; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[g][j] = 0.f;
    if (NM > 0) {
      for (int i = 0; i < n_my; ++i) {
        const int st = i % G::kStages;
        mbar_wait(&full[st], (i / G::kStages) & 1);
        if (a.db != nullptr) {
          const unsigned char* tile = ring + (size_t)st * G::kStageBytes;
#pragma unroll
          for (int g = 0; g < kGw; ++g) {
            const uint4 c = *reinterpret_cast<const uint4*>(tile + (warp * kGw + g) * 512 + lane * 16);
            const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
              acc[g][2 * j] += f.x; acc[g][2 * j + 1] += f.y;
            }
          }
        }
        mbar_arrive(&empty[st]);
      }
    }
    const float inv = (NM > 0 || !kHead) ? 1.0f / __ldg(a.scale) : 1.0f;
    if (NM > 0 && a.db != nullptr && n_my > 0) {
#pragma unroll
      for (int g = 0; g < kGw; ++g)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v = acc[g][j];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if (lane == ((g * 8 + j) & 31)) atomicAdd(a.db + (warp * kGw + g) * 8 + j, v * inv);
        }
    }
    // ======================= epilogue: TMEM -> smem -> scaled atomics =======================
    // (the staging buffer aliases the ring: every warp must be done reading tiles before anyone writes it)
    asm volatile("bar.sync 1, %0;" ::"n"(kW16EpiWarps * 32) : "memory");
    if (n_my > 0) {
      mbar_wait(&d_full, 0);
      tc_fence_after();
      float* out = reinterpret_cast<float*>(ring);        // [128][FB + 4]; every copy and MMA has retired
      constexpr int kLd = G::kOutLd;
      constexpr int kBlocks = NM + (kHead ? 1 : 0);
#pragma unroll 1
      for (int mb = 0; mb < kBlocks; ++mb) {
        const bool head = kHead && mb == NM;
        const float sc = head ? 1.0f / __ldg(a.scale2) : inv;
        const int rows = head ? 8 : 128;                  // head block: only its first 8 rows are distinct
        if (warp * 32 < rows) {
          const int row = warp * 32 + lane;
#pragma unroll 1
          for (int c0 = 0; c0 < FB; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + mb * FB + c0, v);
            tmem_wait_ld();
            if (row < rows) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(out + row * kLd + c0 + j) =
                    make_float4(__uint_as_float(v[j]) * sc, __uint_as_float(v[j + 1]) * sc, __uint_as_float(v[j + 2]) * sc,
                                __uint_as_float(v[j + 3]) * sc);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kW16EpiWarps * 32) : "memory");
        if (head) {
          // features 4..7 of the hg cell are the fp16 residuals of features 0..3: row r + row r + 4
          for (int e = tid; e < 4 * FB; e += kW16EpiWarps * 32) {
            const int r = e / FB, k = e - r * FB;
            if (a.dH[r] != nullptr && k < a.K) atomicAdd(a.dH[r] + k, out[r * kLd + k] + out[(r + 4) * kLd + k]);
          }
        } else if (a.K == FB && ((a.ldw | a.col_off) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dW) & 15) == 0) {
          for (int e = tid; e < 128 * (FB / 4); e += kW16EpiWarps * 32) {
            const int m = e / (FB / 4), k = (e - m * (FB / 4)) * 4;
            const float4 v = *reinterpret_cast<const float4*>(out + m * kLd + k);
            float* dst = a.dW + (size_t)(mb * 128 + m) * a.ldw + a.col_off + k;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
          }
        } else {
          for (int e = tid; e < 128 * FB; e += kW16EpiWarps * 32) {
            const int m = e / FB, k = e - m * FB;
            if (k < a.K) atomicAdd(a.dW + (size_t)(mb * 128 + m) * a.ldw + a.col_off + k, out[m * kLd + k]);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kW16EpiWarps * 32) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kW16MmaWarp) tmem_dealloc<G::kTmemCols>(tbase);
}

template <int NM, int FB, bool kHead>
int launch_wgrad16(W16Args a, cudaStream_t st) {
  using G = W16Geo<NM, FB, kHead>;
  static SmemOptIn optin;
  if (int rc = ensure_smem(wgrad16_kernel<NM, FB, kHead>, optin, G::kSmemBytes, "wgrad16")) return rc;
  long long ctas = sm_count();
  if (ctas > a.n_tiles) ctas = a.n_tiles;
  a.tiles_per_cta = (a.n_tiles + ctas - 1) / ctas;
  ctas = (a.n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
  wgrad16_kernel<NM, FB, kHead><<<(unsigned)ctas, kW16Threads, G::kSmemBytes, st>>>(a);
  return check_launch("wgrad16_kernel");
}

}  // namespace

// dY: (Ppad, FA) T32 fp16 scaled by *scale, FA = 128 or 256 (0 with dY == nullptr: head rows only);
// X: (Ppad, FB) T32 fp16, FB in {256, 128, 64, 32}, first K columns valid;
// hg (nullable): (Ppad, 8) T32 fp16 scaled by *scale2 -> dH[r] (r < 8, nullable) += hg[:, r]^T X.
int run_wgrad16(const void* dY, int FA, const void* X, int FB, int K, float* dW, int ldw, int col_off, float* db,
                const float* scale, const void* hg, float* const* dH, const float* scale2, long long n_points_pad,
                cudaStream_t st) {
  if (n_points_pad == 0) return SNB_OK;
  W16Args a{};
  a.dY = reinterpret_cast<const unsigned char*>(dY);
  a.X = reinterpret_cast<const unsigned char*>(X);
  a.hg = reinterpret_cast<const unsigned char*>(hg);
  a.K = K; a.dW = dW; a.ldw = ldw; a.col_off = col_off; a.db = db; a.scale = scale; a.scale2 = scale2;
  for (int r = 0; r < 8; ++r) a.dH[r] = (hg != nullptr && dH != nullptr) ? dH[r] : nullptr;
  a.n_tiles = n_points_pad / kA16Tile;
  const bool head = hg != nullptr;
  if (FA == 256 && FB == 256 && !head) return launch_wgrad16<2, 256, false>(a, st);
  if (FA == 128 && FB == 256 && head) return launch_wgrad16<1, 256, true>(a, st);
  if (FA == 128 && FB == 256 && !head) return launch_wgrad16<1, 256, false>(a, st);
  if (FA == 256 && FB == 64 && !head) return launch_wgrad16<2, 64, false>(a, st);
  if (FA == 128 && FB == 32 && !head) return launch_wgrad16<1, 32, false>(a, st);
  if (FA == 0 && FB == 128 && head) return launch_wgrad16<0, 128, true>(a, st);
  return fail(SNB_ERR_INVALID, "run_wgrad16: unsupported shape FA=%d FB=%d head=%d", FA, FB, (int)head);
}

}  // namespace snb
