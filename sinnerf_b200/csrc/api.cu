// api.cu -- the extern "C" surface of libsinnerf_b200.so (include/sinnerf_b200.h).
// Host-side argument checking and stage sequencing only; kernels live in the other .cu files.
#include <stdarg.h>

#include "common.cuh"

namespace snb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SNB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SNB_OK;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return dev;
}

int sm_count() {
  static int sms[kMaxDevices];
  const int dev = current_device() & (kMaxDevices - 1);
  if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

// launchers defined in ray_kernels.cu / field_simt.cu / field_tc.cu
int launch_sample_coarse(const float*, const float*, const float*, float, int, int64_t, int, float*, cudaStream_t);
int launch_embed(const float*, int64_t, int, int, float*, cudaStream_t);
int launch_composite(const float*, int, const float*, const float*, const float*, float, int, int64_t, int,
                     float*, float*, float*, const SnbLossSpec*, float*, float*, const SnbPixelScatter*, cudaStream_t);
int launch_sample_pdf(const float*, int64_t, const float*, int64_t, const float*, int64_t, int64_t, int, int,
                      float, float*, cudaStream_t);
int launch_importance_merge(const float*, const float*, const float*, int64_t, int64_t, int, int, float, float*,
                            float*, cudaStream_t);
int launch_pack_fp32(const float* const*, int, void*, int, cudaStream_t);
int field_forward_fp32(const void*, const float*, const float*, int64_t, int, int, float*, cudaStream_t);
int mlp_forward_fp32(const void*, const float*, int64_t, int64_t, int, float*, cudaStream_t);
int field_forward_train_fp32(const void*, const float*, const float*, int64_t, int, float*, float*, float*, float*,
                             float*, cudaStream_t);
int launch_composite_bwd(const float*, const float*, const float*, const float*, float, int, const float*,
                         const float*, const float*, int64_t, int, float*, const SnbLossSpec*, const float*,
                         const float*, const float*, float*, cudaStream_t);
int field_backward_fp32(const float* const*, float* const*, int, const float*, const float*, const float*,
                        const float*, const float*, const float*, int64_t, float*, float*, float*, float*,
                        uint32_t*, cudaStream_t);
int launch_generate_rays(const float*, float, float, float, float, float, float, int, int, int, int, int, int, float*,
                         cudaStream_t);
int field_forward_train16_tc(const void*, int, const float*, const float*, int64_t, int, float*, void*, cudaStream_t);
size_t act16_bytes(long long);
size_t bwd16_workspace_bytes(long long);
int field_backward16(const float* const*, float* const*, int, const float*, const float*, const void*, long long, void*,
                     const float*, cudaStream_t);
int adam_step_pack(float* const*, const float* const*, float*, float*, const SnbAdamArgs&, int, int, void*, cudaStream_t);
// tensor-core modes (field_tc.cu)
size_t tc_packed_bytes(int precision);
int launch_pack_tc(const float* const*, int, int, void*, int, cudaStream_t);
int field_forward_tc(const void*, int, const float*, const float*, int64_t, int, int, float*, cudaStream_t);
int mlp_forward_tc(const void*, int, const float*, int64_t, int64_t, int, float*, cudaStream_t);
int field_forward_train_tc(const void*, int, const float*, const float*, int64_t, int, float*, float*, float*, float*,
                           float*, cudaStream_t);

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_precision(int precision) {
  if (precision < SNB_PREC_FP32 || precision > SNB_PREC_BF16)
    return fail(SNB_ERR_INVALID, "unknown precision mode %d", precision);
  return SNB_OK;
}

}  // namespace snb

using namespace snb;

extern "C" {

int snb_version(void) { return SNB_VERSION; }

const char* snb_last_error(void) { return g_last_error.c_str(); }

int snb_device_check(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(SNB_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  if (major != 10)
    return fail(SNB_ERR_UNSUPPORTED, "libsinnerf_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major,
                minor);
  return SNB_OK;
}

size_t snb_packed_weights_bytes(int precision) {
  if (precision == SNB_PREC_FP32) return sizeof(PackedHeader) + sizeof(float) * (size_t)make_fp32_layout().total;
  if (precision >= SNB_PREC_F16X3 && precision <= SNB_PREC_BF16) return tc_packed_bytes(precision);
  return 0;
}

static int pack_weights_impl(const char* who, const float* const* params, int precision, int new_activation,
                             void* packed, int only_if_dirty, void* stream) {
  SNB_REQUIRE(params != nullptr && packed != nullptr, "%s: null pointer", who);
  SNB_REQUIRE(aligned16(packed), "%s: packed image must be 16-byte aligned", who);
  for (int i = 0; i < SNB_N_PARAM_TENSORS; ++i)
    SNB_REQUIRE(params[i] != nullptr, "%s: parameter tensor %d is null", who, i);
  if (int rc = check_precision(precision)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // the check kernel always runs: it also stamps the header with the checksum of what is being packed
  ParamPtrs pp;
  for (int i = 0; i < SNB_N_PARAM_TENSORS; ++i) pp.p[i] = params[i];
  if (int rc = launch_params_check(pp, precision, new_activation ? 1 : 0, packed, st)) return rc;
  if (precision == SNB_PREC_FP32) return launch_pack_fp32(params, new_activation ? 1 : 0, packed, only_if_dirty, st);
  return launch_pack_tc(params, precision, new_activation ? 1 : 0, packed, only_if_dirty, st);
}

int snb_pack_weights(const float* const* params, int precision, int new_activation, void* packed, void* stream) {
  return pack_weights_impl("snb_pack_weights", params, precision, new_activation, packed, 0, stream);
}

int snb_refresh_weights(const float* const* params, int precision, int new_activation, void* packed, void* stream) {
  return pack_weights_impl("snb_refresh_weights", params, precision, new_activation, packed, 1, stream);
}

int snb_sample_coarse(const float* rays, const float* z_steps, const float* perturb_u, float perturb, int use_disp,
                      int64_t n_rays, int n_samples, float* z_vals, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_sample_coarse: bad extents (%lld rays, %d samples)",
              (long long)n_rays, n_samples);
  SNB_REQUIRE(n_rays == 0 || (rays && z_steps && z_vals), "snb_sample_coarse: null pointer");
  SNB_REQUIRE(!(perturb > 0.f) || perturb_u != nullptr, "snb_sample_coarse: perturb > 0 needs perturb_u");
  return launch_sample_coarse(rays, z_steps, perturb_u, perturb, use_disp, n_rays, n_samples, z_vals,
                              reinterpret_cast<cudaStream_t>(stream));
}

int snb_embed(const float* x, int64_t n, int in_channels, int n_freqs, float* out, void* stream) {
  SNB_REQUIRE(n >= 0 && in_channels >= 1 && n_freqs >= 0 && n_freqs <= 24, "snb_embed: bad extents");
  SNB_REQUIRE(n == 0 || (x && out), "snb_embed: null pointer");
  return launch_embed(x, n, in_channels, n_freqs, out, reinterpret_cast<cudaStream_t>(stream));
}

int snb_mlp_forward(const void* packed, int precision, const float* x, int64_t x_stride, int64_t n_points,
                    int sigma_only, float* out, void* stream) {
  SNB_REQUIRE(n_points >= 0, "snb_mlp_forward: negative point count");
  SNB_REQUIRE(n_points == 0 || (packed && x && out), "snb_mlp_forward: null pointer");
  SNB_REQUIRE(x_stride >= (sigma_only ? SNB_XYZ_CH : SNB_XYZ_CH + SNB_DIR_CH),
              "snb_mlp_forward: row stride %lld shorter than the embedded row", (long long)x_stride);
  SNB_REQUIRE(sigma_only || aligned16(out), "snb_mlp_forward: (P,4) output must be 16-byte aligned");
  if (int rc = check_precision(precision)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == SNB_PREC_FP32) return mlp_forward_fp32(packed, x, x_stride, n_points, sigma_only, out, st);
  return mlp_forward_tc(packed, precision, x, x_stride, n_points, sigma_only, out, st);
}

int snb_field_forward(const void* packed, int precision, const float* rays, const float* z_vals, int64_t n_rays,
                      int n_samples, int sigma_only, float* raw, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_field_forward: bad extents");
  SNB_REQUIRE(n_rays == 0 || (packed && rays && z_vals && raw), "snb_field_forward: null pointer");
  SNB_REQUIRE(aligned16(rays), "snb_field_forward: rays must be 16-byte aligned");
  SNB_REQUIRE(sigma_only || aligned16(raw), "snb_field_forward: (N,S,4) output must be 16-byte aligned");
  if (int rc = check_precision(precision)) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == SNB_PREC_FP32) return field_forward_fp32(packed, rays, z_vals, n_rays, n_samples, sigma_only, raw, st);
  return field_forward_tc(packed, precision, rays, z_vals, n_rays, n_samples, sigma_only, raw, st);
}

int snb_composite_forward(const float* raw, int raw_channels, const float* z_vals, const float* rays,
                          const float* noise, float noise_std, int white_back, int64_t n_rays, int n_samples,
                          float* rgb, float* depth, float* weights, void* stream) {
  SNB_REQUIRE(raw_channels == 4 || raw_channels == 1, "snb_composite_forward: raw_channels must be 4 or 1");
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_composite_forward: bad extents");
  SNB_REQUIRE(n_rays == 0 || (raw && z_vals && rays && weights), "snb_composite_forward: null pointer");
  SNB_REQUIRE(n_rays == 0 || raw_channels == 1 || (rgb && depth), "snb_composite_forward: rgb/depth outputs required");
  SNB_REQUIRE(raw_channels == 1 || aligned16(raw), "snb_composite_forward: raw must be 16-byte aligned");
  const float* nz = (noise_std != 0.f) ? noise : nullptr;
  return launch_composite(raw, raw_channels, z_vals, rays, nz, noise_std, white_back, n_rays, n_samples, rgb,
                          depth, weights, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

int snb_composite_forward_scatter(const float* raw, const float* z_vals, const float* rays, const float* noise,
                                  float noise_std, int white_back, int64_t n_rays, int n_samples, float* rgb,
                                  float* depth, float* weights, const SnbPixelScatter* scatter, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_composite_forward_scatter: bad extents");
  SNB_REQUIRE(n_rays == 0 || (raw && z_vals && rays && weights && rgb && depth), "snb_composite_forward_scatter: null pointer");
  SNB_REQUIRE(aligned16(raw), "snb_composite_forward_scatter: raw must be 16-byte aligned");
  SNB_REQUIRE(scatter != nullptr && scatter->n_dst >= 1 && scatter->n_dst <= SNB_MAX_PIXEL_DST && scatter->row_offset >= 0,
              "snb_composite_forward_scatter: scatter needs 1..%d destinations and a non-negative row offset", SNB_MAX_PIXEL_DST);
  for (int i = 0; i < scatter->n_dst; ++i)
    SNB_REQUIRE(scatter->dst[i] != nullptr && aligned16(scatter->dst[i]),
                "snb_composite_forward_scatter: destination %d is null or not 16-byte aligned", i);
  const float* nz = (noise_std != 0.f) ? noise : nullptr;
  return launch_composite(raw, 4, z_vals, rays, nz, noise_std, white_back, n_rays, n_samples, rgb, depth, weights,
                          nullptr, nullptr, nullptr, scatter, reinterpret_cast<cudaStream_t>(stream));
}

static int check_loss_spec(const char* who, const SnbLossSpec* loss) {
  SNB_REQUIRE(loss != nullptr, "%s: null loss spec", who);
  SNB_REQUIRE(loss->target_rgb != nullptr || loss->target_depth != nullptr, "%s: the loss spec has no target", who);
  return SNB_OK;
}

int snb_composite_forward_loss(const float* raw, const float* z_vals, const float* rays, const float* noise,
                               float noise_std, int white_back, int64_t n_rays, int n_samples,
                               const SnbLossSpec* loss, float* rgb, float* depth, float* weights, float* loss_out,
                               float* loss_ws, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_composite_forward_loss: bad extents");
  if (int rc = check_loss_spec("snb_composite_forward_loss", loss)) return rc;
  SNB_REQUIRE(loss_out != nullptr && loss_ws != nullptr, "snb_composite_forward_loss: null loss output / workspace");
  SNB_REQUIRE(n_rays == 0 || (raw && z_vals && rays && weights && rgb && depth), "snb_composite_forward_loss: null pointer");
  SNB_REQUIRE(aligned16(raw), "snb_composite_forward_loss: raw must be 16-byte aligned");
  const float* nz = (noise_std != 0.f) ? noise : nullptr;
  return launch_composite(raw, 4, z_vals, rays, nz, noise_std, white_back, n_rays, n_samples, rgb, depth, weights,
                          loss, loss_out, loss_ws, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

int snb_sample_pdf(const float* bins, int64_t bins_stride, const float* weights, int64_t w_stride, const float* u,
                   int64_t u_stride, int64_t n_rays, int m, int n_importance, float eps, float* samples,
                   void* stream) {
  SNB_REQUIRE(n_rays >= 0 && m >= 1 && n_importance >= 1, "snb_sample_pdf: bad extents");
  SNB_REQUIRE(n_rays == 0 || (bins && weights && u && samples), "snb_sample_pdf: null pointer");
  SNB_REQUIRE(bins_stride >= m + 1 && w_stride >= m, "snb_sample_pdf: row strides shorter than rows");
  SNB_REQUIRE(u_stride == 0 || u_stride >= n_importance, "snb_sample_pdf: bad u stride");
  return launch_sample_pdf(bins, bins_stride, weights, w_stride, u, u_stride, n_rays, m, n_importance, eps,
                           samples, reinterpret_cast<cudaStream_t>(stream));
}

int snb_importance_merge(const float* z_coarse, const float* weights_coarse, const float* u, int64_t u_stride,
                         int64_t n_rays, int n_samples, int n_importance, float eps, float* z_fine, float* z_new,
                         void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 3 && n_importance >= 1,
              "snb_importance_merge: needs N_samples >= 3 and N_importance >= 1");
  SNB_REQUIRE(n_rays == 0 || (z_coarse && weights_coarse && u && z_fine), "snb_importance_merge: null pointer");
  SNB_REQUIRE(u_stride == 0 || u_stride >= n_importance, "snb_importance_merge: bad u stride");
  return launch_importance_merge(z_coarse, weights_coarse, u, u_stride, n_rays, n_samples, n_importance, eps,
                                 z_fine, z_new, reinterpret_cast<cudaStream_t>(stream));
}

int snb_generate_rays(const float* c2w, float fx, float fy, float cx, float cy, float near, float far, int opencv,
                      int row0, int col0, int rows, int cols, int stride, float* rays, void* stream) {
  SNB_REQUIRE(c2w != nullptr, "snb_generate_rays: null camera matrix");
  SNB_REQUIRE(rows >= 0 && cols >= 0 && stride >= 1 && row0 >= 0 && col0 >= 0, "snb_generate_rays: bad window");
  SNB_REQUIRE(fx != 0.f && fy != 0.f, "snb_generate_rays: zero focal length");
  SNB_REQUIRE((long long)rows * cols == 0 || (rays != nullptr && aligned16(rays)),
              "snb_generate_rays: rays must be a 16-byte aligned device buffer");
  return launch_generate_rays(c2w, fx, fy, cx, cy, near, far, opencv, row0, col0, rows, cols, stride, rays,
                              reinterpret_cast<cudaStream_t>(stream));
}

int snb_field_forward_train(const void* packed, int precision, const float* rays, const float* z_vals,
                            int64_t n_rays, int n_samples, float* raw, float* save_enc, float* save_dir,
                            float* save_h, float* save_g, void* stream) {
  if (int rc = check_precision(precision)) return rc;
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_field_forward_train: bad extents");
  SNB_REQUIRE(n_rays == 0 || (packed && rays && z_vals && raw && save_enc && save_dir && save_h && save_g),
              "snb_field_forward_train: null pointer");
  SNB_REQUIRE(aligned16(rays) && aligned16(raw) && aligned16(save_enc) && aligned16(save_dir) && aligned16(save_h) &&
                  aligned16(save_g),
              "snb_field_forward_train: buffers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == SNB_PREC_FP32)
    return field_forward_train_fp32(packed, rays, z_vals, n_rays, n_samples, raw, save_enc, save_dir, save_h, save_g, st);
  return field_forward_train_tc(packed, precision, rays, z_vals, n_rays, n_samples, raw, save_enc, save_dir, save_h,
                                save_g, st);
}

int snb_composite_backward(const float* raw, const float* z_vals, const float* rays, const float* noise,
                           float noise_std, int white_back, const float* g_rgb, const float* g_depth,
                           const float* g_weights, int64_t n_rays, int n_samples, float* g_raw, void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_composite_backward: bad extents");
  SNB_REQUIRE(n_rays == 0 || (raw && z_vals && rays && g_raw), "snb_composite_backward: null pointer");
  SNB_REQUIRE(aligned16(raw) && aligned16(g_raw), "snb_composite_backward: raw / g_raw must be 16-byte aligned");
  const float* nz = (noise_std != 0.f) ? noise : nullptr;
  return launch_composite_bwd(raw, z_vals, rays, nz, noise_std, white_back, g_rgb, g_depth, g_weights, n_rays,
                              n_samples, g_raw, nullptr, nullptr, nullptr, nullptr, nullptr,
                              reinterpret_cast<cudaStream_t>(stream));
}

int snb_composite_backward_loss(const float* raw, const float* z_vals, const float* rays, const float* noise,
                                float noise_std, int white_back, const float* g_rgb, const float* g_depth,
                                const float* g_weights, const SnbLossSpec* loss, const float* rgb, const float* depth,
                                const float* g_loss, int64_t n_rays, int n_samples, float* g_raw, float* g_amax,
                                void* stream) {
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_composite_backward_loss: bad extents");
  SNB_REQUIRE(n_rays == 0 || (raw && z_vals && rays && g_raw), "snb_composite_backward_loss: null pointer");
  SNB_REQUIRE(aligned16(raw) && aligned16(g_raw), "snb_composite_backward_loss: raw / g_raw must be 16-byte aligned");
  if (loss != nullptr) {
    if (int rc = check_loss_spec("snb_composite_backward_loss", loss)) return rc;
    SNB_REQUIRE(n_rays == 0 || ((loss->target_rgb == nullptr || rgb) && (loss->target_depth == nullptr || depth)),
                "snb_composite_backward_loss: the forward's rgb / depth outputs are required with a loss spec");
  }
  const float* nz = (noise_std != 0.f) ? noise : nullptr;
  return launch_composite_bwd(raw, z_vals, rays, nz, noise_std, white_back, g_rgb, g_depth, g_weights, n_rays,
                              n_samples, g_raw, loss, rgb, depth, g_loss, g_amax, reinterpret_cast<cudaStream_t>(stream));
}

int snb_field_backward(const float* const* params, float* const* grads, int new_activation, const float* g_raw,
                       const float* raw, const float* save_enc, const float* save_dir, const float* save_h,
                       const float* save_g, int64_t n_points, float* ws_a, float* ws_b, float* ws_s,
                       float* ws_w, uint32_t* ws_m, void* stream) {
  SNB_REQUIRE(n_points >= 0, "snb_field_backward: negative point count");
  SNB_REQUIRE(params && grads, "snb_field_backward: null parameter arrays");
  for (int i = 0; i < SNB_N_PARAM_TENSORS; ++i)
    SNB_REQUIRE(params[i] && grads[i], "snb_field_backward: parameter / gradient tensor %d is null", i);
  SNB_REQUIRE(n_points == 0 || (g_raw && raw && save_enc && save_dir && save_h && save_g && ws_a && ws_b && ws_s && ws_w && ws_m),
              "snb_field_backward: null pointer");
  return field_backward_fp32(params, grads, new_activation, g_raw, raw, save_enc, save_dir, save_h, save_g,
                             n_points, ws_a, ws_b, ws_s, ws_w, ws_m, reinterpret_cast<cudaStream_t>(stream));
}

size_t snb_act16_bytes(int64_t n_points) { return n_points < 0 ? 0 : act16_bytes(n_points); }
size_t snb_bwd16_workspace_bytes(int64_t n_points) { return n_points < 0 ? 0 : bwd16_workspace_bytes(n_points); }

int snb_field_forward_train16(const void* packed, int precision, const float* rays, const float* z_vals, int64_t n_rays,
                              int n_samples, float* raw, void* act16, void* stream) {
  if (int rc = check_precision(precision)) return rc;
  if (precision == SNB_PREC_FP32)
    return fail(SNB_ERR_UNSUPPORTED, "snb_field_forward_train16: 16-bit activation storage needs a tensor-core precision mode");
  SNB_REQUIRE(n_rays >= 0 && n_samples >= 1, "snb_field_forward_train16: bad extents");
  SNB_REQUIRE(n_rays == 0 || (packed && rays && z_vals && raw && act16), "snb_field_forward_train16: null pointer");
  SNB_REQUIRE(aligned16(rays) && aligned16(raw) && (reinterpret_cast<uintptr_t>(act16) & 255u) == 0,
              "snb_field_forward_train16: rays / raw must be 16-byte and act16 256-byte aligned");
  return field_forward_train16_tc(packed, precision, rays, z_vals, n_rays, n_samples, raw, act16,
                                  reinterpret_cast<cudaStream_t>(stream));
}

int snb_field_backward16(const float* const* params, float* const* grads, int new_activation, const float* g_raw,
                         const float* raw, const void* act16, int64_t n_points, void* workspace, const float* g_amax,
                         void* stream) {
  SNB_REQUIRE(n_points >= 0, "snb_field_backward16: negative point count");
  SNB_REQUIRE(params && grads, "snb_field_backward16: null parameter arrays");
  for (int i = 0; i < SNB_N_PARAM_TENSORS; ++i)
    SNB_REQUIRE(params[i] && grads[i], "snb_field_backward16: parameter / gradient tensor %d is null", i);
  SNB_REQUIRE(n_points == 0 || (g_raw && raw && act16 && workspace), "snb_field_backward16: null pointer");
  SNB_REQUIRE(aligned16(g_raw) && aligned16(raw) && (reinterpret_cast<uintptr_t>(act16) & 255u) == 0 &&
                  (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
              "snb_field_backward16: g_raw / raw must be 16-byte, act16 / workspace 256-byte aligned");
  return field_backward16(params, grads, new_activation, g_raw, raw, act16, n_points, workspace, g_amax,
                          reinterpret_cast<cudaStream_t>(stream));
}

int snb_adam_step(float* const* params, const float* const* grads, float* exp_avg, float* exp_avg_sq,
                  const SnbAdamArgs* args, int precision, int new_activation, void* packed, void* stream) {
  SNB_REQUIRE(params && grads && exp_avg && exp_avg_sq && args, "snb_adam_step: null pointer");
  for (int i = 0; i < SNB_N_PARAM_TENSORS; ++i) SNB_REQUIRE(params[i] != nullptr, "snb_adam_step: parameter tensor %d is null", i);
  SNB_REQUIRE(args->step >= 1, "snb_adam_step: step counts from 1 (got %d)", args->step);
  SNB_REQUIRE(args->lr >= 0. && args->eps >= 0. && args->beta1 >= 0. && args->beta1 < 1. && args->beta2 >= 0. &&
                  args->beta2 < 1. && args->weight_decay >= 0.,
              "snb_adam_step: invalid hyper-parameters");
  SNB_REQUIRE(packed == nullptr || aligned16(packed), "snb_adam_step: packed image must be 16-byte aligned");
  if (packed != nullptr)
    if (int rc = check_precision(precision)) return rc;
  static_assert(SNB_PARAM_FLOATS == 593408 + 2436, "parameter count");
  return adam_step_pack(params, grads, exp_avg, exp_avg_sq, *args, precision, new_activation, packed,
                        reinterpret_cast<cudaStream_t>(stream));
}

int snb_render_forward(const SnbRenderArgs* a, void* stream) {
  SNB_REQUIRE(a != nullptr, "snb_render_forward: null args");
  SNB_REQUIRE(a->n_rays >= 0 && a->n_samples >= 1 && a->n_importance >= 0, "snb_render_forward: bad extents");
  if (a->n_rays == 0) return SNB_OK;
  SNB_REQUIRE(a->rays && a->packed_coarse && a->z_steps && a->z_coarse && a->raw_coarse && a->weights_coarse,
              "snb_render_forward: null coarse-pass pointer");
  SNB_REQUIRE(a->test_time || (a->rgb_coarse && a->depth_coarse), "snb_render_forward: coarse outputs required");
  // rendering.py:330-333 dereferences rgb_coarse, which test_time never defines
  SNB_REQUIRE(!(a->test_time && a->n_importance == 0),
              "snb_render_forward: test_time requires N_importance > 0 (the reference raises UnboundLocalError)");
  const int S = a->n_samples, Ni = a->n_importance;
  int rc;
  if ((rc = snb_sample_coarse(a->rays, a->z_steps, a->perturb_u, a->perturb, a->use_disp, a->n_rays, S,
                              a->z_coarse, stream)))
    return rc;
  if ((rc = snb_field_forward(a->packed_coarse, a->precision, a->rays, a->z_coarse, a->n_rays, S, a->test_time,
                              a->raw_coarse, stream)))
    return rc;
  if (Ni == 0 && a->pixel_scatter != nullptr)     // the coarse pass is the last one: its pixels are the frame's
    return snb_composite_forward_scatter(a->raw_coarse, a->z_coarse, a->rays, a->noise_coarse, a->noise_std, a->white_back,
                                         a->n_rays, S, a->rgb_coarse, a->depth_coarse, a->weights_coarse, a->pixel_scatter,
                                         stream);
  if ((rc = snb_composite_forward(a->raw_coarse, a->test_time ? 1 : 4, a->z_coarse, a->rays, a->noise_coarse,
                                  a->noise_std, a->white_back, a->n_rays, S, a->rgb_coarse, a->depth_coarse,
                                  a->weights_coarse, stream)))
    return rc;
  if (Ni == 0) return SNB_OK;
  SNB_REQUIRE(a->packed_fine && a->z_fine && a->raw_fine && a->rgb_fine && a->depth_fine && a->weights_fine,
              "snb_render_forward: null fine-pass pointer");
  const bool det = !(a->perturb > 0.f);
  const float* u = det ? a->u_steps : a->pdf_u;
  SNB_REQUIRE(u != nullptr, "snb_render_forward: %s required", det ? "u_steps" : "pdf_u");
  if ((rc = snb_importance_merge(a->z_coarse, a->weights_coarse, u, det ? 0 : Ni, a->n_rays, S, Ni, 1e-5f,
                                 a->z_fine, nullptr, stream)))
    return rc;
  if ((rc = snb_field_forward(a->packed_fine, a->precision, a->rays, a->z_fine, a->n_rays, S + Ni, 0, a->raw_fine,
                              stream)))
    return rc;
  if (a->pixel_scatter != nullptr)
    return snb_composite_forward_scatter(a->raw_fine, a->z_fine, a->rays, a->noise_fine, a->noise_std, a->white_back,
                                         a->n_rays, S + Ni, a->rgb_fine, a->depth_fine, a->weights_fine, a->pixel_scatter,
                                         stream);
  return snb_composite_forward(a->raw_fine, 4, a->z_fine, a->rays, a->noise_fine, a->noise_std, a->white_back,
                               a->n_rays, S + Ni, a->rgb_fine, a->depth_fine, a->weights_fine, stream);
}

}  // extern "C"
